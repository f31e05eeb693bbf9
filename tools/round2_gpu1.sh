#!/usr/bin/env bash
# round 2, first GPU call: correctness of the new paths + first timings
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r02a_pytest.log
tail -3 $O/r02a_pytest.log
timeout 600 python bench.py > $O/r02a_bench.json 2> $O/r02a_bench.err
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d.get('e2e',{}).get('ms_per_step',0),2))"; }
cat $O/r02a_bench.json | show default
SMART_B200_FAST_REGS=lean  timeout 200 python bench.py --no-also --no-cpu-baseline --steps 5 | tee $O/r02a_c2_lean.json | show c2-lean
SMART_B200_FAST_REGS=roomy timeout 200 python bench.py --no-also --no-cpu-baseline --steps 5 | tee $O/r02a_c2_roomy.json | show c2-roomy
timeout 300 python bench.py --workload c3 --members 600000 --no-also --no-cpu-baseline --no-e2e --steps 2 | tee $O/r02a_c3_600k.json | show c3-600k
timeout 200 python bench.py --flags 65536 --no-also --steps 3 --no-cpu-baseline | tee $O/r02a_c2_perstep.json | show c2-perstep
timeout 200 python bench.py --flags 1 --no-also --steps 3 --no-cpu-baseline | tee $O/r02a_c2_general.json | show c2-general
timeout 300 python tools/wild_members_bench.py > $O/r02a_wild.json 2> $O/r02a_wild.err; cat $O/r02a_wild.json
SMART_B200_NO_SIDE_STREAM=1 timeout 300 python tools/wild_members_bench.py > $O/r02a_wild_serial.json 2>> $O/r02a_wild.err; cat $O/r02a_wild_serial.json
ls -la $O | tail -20
