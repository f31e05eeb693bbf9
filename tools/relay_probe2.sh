#!/usr/bin/env bash
set -u
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 4"
export SMART_B200_RELAY_MIN_WAVES=0.2 SMART_B200_RELAY_MAX_WAVES=1000
$B --workload c5 | show c5-relay-minwaves0.2
for s in 16 64; do SMART_B200_RELAY_SEGS=$s $B --workload c5 | show c5-segs$s; done
for n in 30000 50000 65000; do $B --members $n | show c2-$n-relay; SMART_B200_RELAY=0 $B --members $n | show c2-$n-norelay; done
SMART_B200_RELAY=0 $B --members 1000000 | show c2-1m-norelay
$B --members 1000000 | show c2-1m-relay-auto
SMART_B200_RELAY_SEGS=8 $B --members 1000000 | show c2-1m-relay-segs8
SMART_B200_RELAY=0 $B --workload c3 --steps 2 | show c3-norelay
$B --workload c3 --steps 2 | show c3-relay-auto
SMART_B200_RELAY_SEGS=8 $B --workload c3 --steps 2 | show c3-relay-segs8
SMART_B200_RELAY_SEGS=16 $B --workload c3 --steps 2 | show c3-relay-segs16
SMART_B200_BLOCK=128 $B | show c2-block128-relay
SMART_B200_FAST_REGS=lean $B | show c2-lean-relay
SMART_B200_RELAY=0 $B --workload c5 --members 1000000 | show c5-1m-norelay
$B --workload c5 --members 1000000 | show c5-1m-relay
