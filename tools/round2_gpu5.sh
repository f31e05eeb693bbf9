#!/usr/bin/env bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r02e_pytest.log; tail -8 $O/r02e_pytest.log
K='regex:smart_batch_kernel<(double|float), \(int\)0'
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -o $O/prof_r02e_c2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02e_c2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -o $O/prof_r02e_c3 \
    python bench.py --workload c3 --members 600000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02e_c3.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -o $O/prof_r02e_c4a \
    python bench.py --workload c4a --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02e_c4a.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 3 -c 1 -o $O/prof_r02e_c5 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02e_c5.log 2>&1
tail -3 $O/ncu_r02e_c2.log | cut -c1-300
ls -la $O/prof_r02e*
