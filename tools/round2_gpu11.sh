#!/usr/bin/env bash
set -u
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 5"
B3="python bench.py --workload c3 --members 600000 --no-also --no-cpu-baseline --no-e2e --steps 2"
SMART_B200_LIB=$PWD/build_exp/lib_park64.so SMART_B200_FAST_REGS=lean timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "members_match or single_run or thirty_year" 2>&1 | tail -3
$B | show c2-base
for v in park64 park72 park80; do
  SMART_B200_LIB=$PWD/build_exp/lib_$v.so SMART_B200_FAST_REGS=lean $B | show c2-$v-lean
done
SMART_B200_LIB=$PWD/build_exp/lib_park80.so SMART_B200_FAST_REGS=roomy $B | show c2-park-roomy96
$B3 | show c3-base
for v in park64 park72 park80; do
  SMART_B200_LIB=$PWD/build_exp/lib_$v.so SMART_B200_FAST_REGS=lean $B3 | show c3-$v-lean
done
SMART_B200_LIB=$PWD/build_exp/lib_park80.so SMART_B200_FAST_REGS=roomy $B3 | show c3-park-roomy96
python bench.py --workload c2 --members 1000000 --no-also --no-cpu-baseline --no-e2e --steps 3 | show c2-1m-base
SMART_B200_LIB=$PWD/build_exp/lib_park64.so SMART_B200_FAST_REGS=lean python bench.py --workload c2 --members 1000000 --no-also --no-cpu-baseline --no-e2e --steps 3 | show c2-1m-park64
