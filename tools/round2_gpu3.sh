#!/usr/bin/env bash
set -u
O=gpurun_out
mkdir -p $O
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), 'G/s', round(d['ms_per_step'],3), 'ms')"; }
B="python bench.py --no-also --no-cpu-baseline --no-e2e --steps 5"
timeout 1800 python -m pytest tests/test_gpu_round2.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 > $O/r02c_pytest.log; tail -5 $O/r02c_pytest.log
echo "== C2 variants" | tee $O/r02c_variants.txt
$B | show base | tee -a $O/r02c_variants.txt
for s in 20000 100000 500000 2000000; do SMART_B200_SKEW=$s $B | show skew$s | tee -a $O/r02c_variants.txt; done
SMART_B200_SKEW=100000 SMART_B200_SKEW_GROUPS=4 $B | show skew100k-g4 | tee -a $O/r02c_variants.txt
for v in lean88 lean72 tree unroll2 unroll8 roomy104; do
  SMART_B200_LIB=$PWD/build_exp/lib_$v.so SMART_B200_FAST_REGS=lean $B | show $v-lean | tee -a $O/r02c_variants.txt
  SMART_B200_LIB=$PWD/build_exp/lib_$v.so SMART_B200_FAST_REGS=roomy $B | show $v-roomy | tee -a $O/r02c_variants.txt
done
echo "== C3 600k variants" | tee -a $O/r02c_variants.txt
B3="python bench.py --workload c3 --members 600000 --no-also --no-cpu-baseline --no-e2e --steps 2"
$B3 | show c3-base | tee -a $O/r02c_variants.txt
SMART_B200_FAST_REGS=lean $B3 | show c3-lean80 | tee -a $O/r02c_variants.txt
for v in lean88 tree unroll2 unroll8 roomy104; do
  SMART_B200_LIB=$PWD/build_exp/lib_$v.so SMART_B200_FAST_REGS=$( [ $v = lean88 ] && echo lean || echo roomy ) $B3 | show c3-$v | tee -a $O/r02c_variants.txt
done
timeout 200 python bench.py --workload c5 --no-also --no-cpu-baseline --steps 5 | show c5 | tee -a $O/r02c_variants.txt
