#!/usr/bin/env bash
# Relay tuning on a B200 box: parity tests of the relay, then C2 / C5 / C2-per-step with the relay off,
# chosen by the library, and with forced segment counts.
set -u
O=gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms', 'e2e', round(d['e2e']['value']/1e9,1) if d.get('e2e') else '')"; }
timeout 900 python -m pytest tests/test_gpu_relay.py -m gpu -x -q 2>&1 | tail -5 | tee $O/relay_pytest.log
B="python bench.py --no-also --no-cpu-baseline --steps 5"
SMART_B200_RELAY=0 $B | show c2-norelay
$B | show c2-auto
for s in 8 16 32 64; do SMART_B200_RELAY_SEGS=$s $B | show c2-segs$s; done
SMART_B200_RELAY=0 $B --workload c5 | show c5-norelay
$B --workload c5 | show c5-auto
SMART_B200_RELAY=0 $B --flags 65536 | show perstep-norelay
$B --flags 65536 | show perstep-auto
for n in 50000 80000 150000 200000 400000; do
  SMART_B200_RELAY=0 $B --members $n --no-e2e | show c2-$n-norelay
  $B --members $n --no-e2e | show c2-$n-auto
done
