#!/usr/bin/env bash
set -u
O=gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r02j_pytest.log
python bench.py --steps 5 > $O/r02j_bench.json 2> $O/r02j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench.json'))
print('c2', round(d['value']/1e9,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e9,1), 'launches', d['gpu_launches'])
for k,v in d['also'].items(): print(k, round(v['value']/1e9,1), round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value']/1e9,1))
PY
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 4"
$B --flags 65536 | show perstep
$B --flags 1 | show general
$B --members 1000000 | show c2-1m
$B --workload c5 --members 1000000 | show c5-1m
