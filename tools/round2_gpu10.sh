#!/usr/bin/env bash
set -u
O=gpurun_out
mkdir -p $O
compute-sanitizer --tool memcheck python tools/sanitizer_cases.py > $O/r02_sanitizer_memcheck.txt 2>&1; tail -4 $O/r02_sanitizer_memcheck.txt
compute-sanitizer --tool racecheck python tools/sanitizer_cases.py > $O/r02_sanitizer_racecheck.txt 2>&1; tail -4 $O/r02_sanitizer_racecheck.txt
python tests/parity_report.py > $O/r02_parity_report.txt 2>&1; cat $O/r02_parity_report.txt
python tools/wild_members_bench.py > $O/r02_wild_members.json 2>&1; cat $O/r02_wild_members.json
SMART_B200_NO_SIDE_STREAM=1 python tools/wild_members_bench.py > $O/r02_wild_members_serial.json 2>&1; cat $O/r02_wild_members_serial.json
