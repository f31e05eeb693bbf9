#!/usr/bin/env bash
# Mnemonic counts and kernel list of the shipped library (profiles/rNN_sass_evidence.txt):
#   bash tools/sass_evidence.sh > profiles/r01_sass_evidence.txt
LIB=${1:-smartpy_b200/libsmart_b200.so}
SASS=$(mktemp)
cuobjdump -sass "$LIB" > "$SASS"
echo "cuobjdump -sass $LIB (sm_100a) -- instruction mnemonic counts over all kernels"
for m in UBLKCP SYNCS.ARRIVE.TRANS64 SYNCS.PHASECHK LDGSTS DFMA DADD DMUL DSETP FFMA VOTE.ANY SHFL BAR.SYNC BAR.RED ATOMS ATOMG RED LDL STL; do
    printf "%-24s %s\n" "$m" "$(grep -c "[[:space:]]$m" "$SASS")"
done
echo
echo "kernels:"
grep "Function :" "$SASS" | awk '{print $3}' | c++filt | sed 's/(anonymous namespace):://g; s/smart:://g' | sort | uniq -c | sort -k2
rm -f "$SASS"
