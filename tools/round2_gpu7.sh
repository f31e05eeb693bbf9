#!/usr/bin/env bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r02g_pytest.log; tail -6 $O/r02g_pytest.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 3"
$B --workload c4a | show c4a-warp-ctas-104
SMART_B200_NO_WARP_CTAS=1 $B --workload c4a | show c4a-old-layout
for v in multi96 multi112 multi128; do SMART_B200_LIB=$PWD/build_exp/lib_$v.so $B --workload c4a | show c4a-$v; done
$B --workload c4b | show c4b-warp-ctas
SMART_B200_NO_WARP_CTAS=1 $B --workload c4b | show c4b-old-layout
$B | show c2
$B --workload c5 | show c5
$B --workload c3 --members 600000 | show c3-600k
K='regex:smart_batch_kernel<(double|float), \(int\)0'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -o $O/prof_r02g_c4a \
    python bench.py --workload c4a --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02g_c4a.log 2>&1
