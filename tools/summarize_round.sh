#!/usr/bin/env bash
# Build box, after tools/measure_round.sh came back: copy what it measured into profiles/.
set -u
R=${1:-rXX}
O=gpurun_out
cp $O/${R}_*_ncu_full_summary.txt $O/${R}_*_source_page.csv.gz profiles/
python tools/merge_traffic.py $O/ncu_traffic_${R}.json
cp $O/${R}_c2_launches.csv $O/${R}_conditioning_bench.json $O/${R}_parity_report.txt $O/${R}_smi.csv profiles/ 2>/dev/null
cp $O/${R}_sanitizer_memcheck.txt $O/${R}_sanitizer_racecheck.txt $O/${R}_wild_members.json profiles/ 2>/dev/null
cp $O/pytest_gpu_${R}.log profiles/${R}_pytest_gpu.log 2>/dev/null
cp $O/fp32_vs_fp64_c2.json profiles/${R}_fp32_vs_fp64_c2.json 2>/dev/null   # written by tests/test_gpu_round2.py
cp $O/fp32_vs_fp64_c3.json profiles/${R}_fp32_vs_fp64_c3.json 2>/dev/null
for f in $O/bench_${R}_*.json; do cp $f profiles/; done
grep -h "pipe_fp64_cycles_active\|gpu__time_duration" profiles/${R}_*_ncu_full_summary.txt | head -40
