#!/usr/bin/env bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r02f_pytest.log; tail -6 $O/r02f_pytest.log
timeout 900 python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err; tail -3 $O/r02f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print('c2', round(d['value']/1e9,1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))
for k,v in d.get('also',{}).items():
    print(k, round(v.get('value',0)/1e9,1), round(v.get('ms_per_step',0),3), 'e2e', round(v.get('e2e',{}).get('ms_per_step',0),3), v.get('error',''))
PY
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
SMART_B200_NO_WARP_PADDING=1 python bench.py --workload c4a --no-also --no-cpu-baseline --no-e2e --steps 3 | show c4a-unpadded
python bench.py --workload c4a --no-also --no-cpu-baseline --no-e2e --steps 3 | show c4a-padded
python bench.py --workload c5 --members 1000000 --no-also --no-cpu-baseline --no-e2e --steps 3 | show c5-1m
python bench.py --workload c2 --members 1000000 --no-also --no-cpu-baseline --no-e2e --steps 3 | show c2-1m
