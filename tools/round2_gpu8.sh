#!/usr/bin/env bash
set -u
O=gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 3"
timeout 900 python -m pytest tests -m gpu -q -k "multi_catchment or c4a or block_sub or ragged" 2>&1 | tail -4
$B --workload c4a | show c4a-warp-ctas-104
for v in multi96 multi112 multi128; do SMART_B200_LIB=$PWD/build_exp/lib_$v.so $B --workload c4a | show c4a-$v; done
$B --workload c4b | show c4b-warp-ctas
for v in multi96 multi128; do SMART_B200_LIB=$PWD/build_exp/lib_$v.so $B --workload c4b | show c4b-$v; done
K='regex:smart_batch_kernel<(double|float), \(int\)0'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -o $O/prof_r02h_c4a \
    python bench.py --workload c4a --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02h_c4a.log 2>&1
