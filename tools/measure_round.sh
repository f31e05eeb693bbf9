#!/usr/bin/env bash
# Reproduces the per-round measurements kept under profiles/ (run on a B200 box, from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/measure_round.sh r01'
# then, back on the build box:  python tools/summarize_profile.py gpurun_out/prof_<round>_c2.ncu-rep 100032 96432
set -u
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu_$R.log
python bench.py > $O/bench_${R}_c2.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${R}_c2_reference_arm.json
# every launch with its device time (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_${R}_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# the dominant kernel, full set, source import (-lineinfo is always on in smartpy_b200/_build.py)
ncu --set full --clock-control none --import-source on -k regex:smart_batch_kernel -s 6 -c 1 -o $O/prof_${R}_c2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_full_$R.log 2>&1
python bench.py --workload c3 --steps 2 --no-cpu-baseline > $O/bench_${R}_c3.json
python bench.py --workload c5 --steps 3 --no-cpu-baseline > $O/bench_${R}_c5.json
python bench.py --workload c4a --steps 2 --no-cpu-baseline > $O/bench_${R}_c4a.json
python bench.py --workload c4b --steps 2 --no-cpu-baseline > $O/bench_${R}_c4b.json
# FP32 mode: the same full capture of its kernel
ncu --set full --clock-control none --import-source on -k regex:smart_batch_kernel -s 6 -c 1 -o $O/prof_${R}_c5 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_full_${R}_c5.log 2>&1
python tools/bench_conditioning.py 10000000 > $O/conditioning_bench_$R.json
python bench.py --flags 65536 --steps 3 --no-cpu-baseline > $O/bench_${R}_c2_perstep.json
python bench.py --flags 1 --steps 3 --no-cpu-baseline > $O/bench_${R}_c2_general.json
python tests/parity_report.py > $O/parity_report_$R.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_$R.csv
ls -la $O
