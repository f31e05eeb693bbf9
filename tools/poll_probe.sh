#!/usr/bin/env bash
set -u
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 5"
timeout 600 python -m pytest tests/test_gpu_relay.py -m gpu -x -q 2>&1 | tail -2
for r in 1 2; do
$B | show c2
$B --workload c5 | show c5
$B --workload c3 --members 600000 --steps 2 | show c3-600k
$B --flags 65536 | show perstep
$B --members 50000 | show c2-50k
done
