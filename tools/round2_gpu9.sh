#!/usr/bin/env bash
set -u
O=gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 3"
timeout 900 python -m pytest tests -m gpu -q -k "multi_catchment or c4a or block_sub or ragged or daily_step" 2>&1 | tail -4
$B --workload c4a | show c4a-96
for v in multi80 multi88 multi128; do SMART_B200_LIB=$PWD/build_exp/lib_$v.so $B --workload c4a | show c4a-$v; done
$B --workload c4b | show c4b-96
for v in multi80 multi88; do SMART_B200_LIB=$PWD/build_exp/lib_$v.so $B --workload c4b | show c4b-$v; done
