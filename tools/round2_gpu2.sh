#!/usr/bin/env bash
set -u
O=gpurun_out
mkdir -p $O
./build_exp/operand_microbench > $O/r02_operand_microbench.txt 2>&1; cat $O/r02_operand_microbench.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r02b_pytest.log
tail -15 $O/r02b_pytest.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d.get('e2e',{}).get('ms_per_step',0),2))"; }
timeout 200 python bench.py --workload c5 --no-also --no-cpu-baseline --steps 5 | tee $O/r02b_c5.json | show c5
timeout 200 python bench.py --workload c5 --members 1000000 --no-also --no-cpu-baseline --steps 3 | tee $O/r02b_c5_1m.json | show c5-1m
timeout 200 python bench.py --workload c2 --members 1000000 --no-also --no-cpu-baseline --steps 3 | tee $O/r02b_c2_1m.json | show c2-1m
ls $O/*.json | head -40
