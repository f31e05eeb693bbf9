#!/usr/bin/env bash
set -u
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 4"
$B --workload c4a --steps 3 | show c4a
$B --workload c4a --precision f32 --steps 3 | show c4a-f32
$B --workload c4b --steps 2 | show c4b
$B | show c2
$B --workload c3 --members 600000 --steps 2 | show c3-600k
