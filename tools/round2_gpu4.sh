#!/usr/bin/env bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "gather_block or run_host" 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:smart_batch_kernel -s 8 -c 1 -o $O/prof_r02d_c2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02d_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:smart_batch_kernel -s 8 -c 1 -o $O/prof_r02d_c3 \
    python bench.py --workload c3 --members 600000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02d_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:smart_batch_kernel -s 2 -c 1 -o $O/prof_r02d_c4a \
    python bench.py --workload c4a --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also > $O/ncu_r02d_c4a.log 2>&1
ls -la $O/*.ncu-rep | tail -5
