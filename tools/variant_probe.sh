#!/usr/bin/env bash
# A/B of kernel variants built into build_exp/lib_<name>.so (tools/build_variant.py): C2 / C3-600k / C4a / C4b / per-step / C5,
# two rounds so that run-to-run noise shows.
set -u
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 4"
for round in 1 2; do
for v in base "$@"; do
  if [ $v = base ]; then unset SMART_B200_LIB; else export SMART_B200_LIB=$PWD/build_exp/lib_$v.so; fi
  $B | show $v-c2
  $B --workload c3 --members 600000 --steps 2 | show $v-c3-600k
  $B --workload c4a --steps 3 | show $v-c4a
  $B --workload c4b --steps 2 | show $v-c4b
  $B --flags 65536 | show $v-perstep
done
done
