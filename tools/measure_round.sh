#!/usr/bin/env bash
# Reproduces the per-round measurements kept under profiles/ (run on a B200 box, from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/measure_round.sh r02'
# then, back on the build box:  bash tools/summarize_round.sh r02   (copies the results into profiles/)
set -u
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu_$R.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${R}_c2_reference_arm.json
# every launch with its device time (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${R}_c2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > /dev/null 2>&1
# the dominant kernel of every configuration, full set, source import (-lineinfo is always on in
# smartpy_b200/_build.py).  The step launches two kernels per run (branch-faithful form first, on
# its own stream): -k picks the fast one (variant 0), -s 3 skips the warm-up runs.
FAST='regex:smart_batch_kernel<(double|float), \(int\)0'
SLOW='regex:smart_batch_kernel<(double|float), \(int\)1'
# Each report (~15 MB) is summarised on the box and dropped (gpurun brings back at most 64 MiB):
# the text summary, the entry of ncu_traffic.json and the per-SASS-line source page stay.
cap() {  # name, kernel filter, key of ncu_traffic.json, members, steps per member, bench arguments
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$2" -s 3 -c 1 -o $O/prof_${R}_$1 \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also ${@:6} > $O/ncu_${R}_$1.log 2>&1
  python tools/summarize_profile.py $O/prof_${R}_$1.ncu-rep $4 $5 $O/${R}_$1_ncu_full_summary.txt \
      --key $3 --json $O/ncu_traffic_${R}.json > /dev/null
  ncu -i $O/prof_${R}_$1.ncu-rep --page source --csv 2> /dev/null | gzip > $O/${R}_$1_source_page.csv.gz
  [ "$1" = c2 ] || rm -f $O/prof_${R}_$1.ncu-rep
}
cap c2 "$FAST" c2:100000:f64:0 100000 96432
cap c3 "$FAST" c3:1250000:f64:0 1250000 271752 --workload c3
cap c4a "$FAST" c4a:1000000:f64:0 1000000 8760 --workload c4a
cap c4b "$FAST" c4b:1000000:f64:0 1000000 87672 --workload c4b
cap c5 "$FAST" c5:100000:f32:0 100000 96432 --workload c5
cap c2_perstep "$FAST" c2:100000:f64:65536 100000 96432 --flags 65536
cap c2_general "$SLOW" c2:100000:f64:1 100000 96432 --flags 1
# the driver's own command, now that captures of the running source exist (roofline.frac_pipe)
python tools/merge_traffic.py $O/ncu_traffic_${R}.json
python bench.py > $O/bench_${R}_c2.json 2> $O/bench_${R}_c2.err
python tools/bench_conditioning.py 10000000 > $O/${R}_conditioning_bench.json
for v in "perstep --flags 65536" "general --flags 1" "1m --members 1000000" "c5_1m --workload c5 --members 1000000"; do
  set -- $v; n=$1; shift
  python bench.py --steps 3 --no-cpu-baseline --no-also "$@" > $O/bench_${R}_c2_$n.json
done
compute-sanitizer --tool memcheck python tools/sanitizer_cases.py 2>&1 | tail -40 > $O/${R}_sanitizer_memcheck.txt
compute-sanitizer --tool racecheck python tools/sanitizer_cases.py 2>&1 | tail -40 > $O/${R}_sanitizer_racecheck.txt
python tools/wild_members_bench.py > $O/${R}_wild_members.json 2>&1
python tests/parity_report.py > $O/${R}_parity_report.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${R}_smi.csv
ls -la $O | tail -40
