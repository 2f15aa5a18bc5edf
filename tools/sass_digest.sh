#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (tcgen05 MMA, TMA, TMEM loads, FP64 tensor MMA)
# in the built library.  Usage: tools/sass_digest.sh > profiles/r02_sass_digest.txt
LIB=${1:-mixmogam_b200/libmixmogam_b200.so}
echo "# $(basename $LIB): SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)"
echo "# UTCIMMA = tcgen05.mma kind::i8, UTCOMMA = tcgen05.mma kind::mxf4.block_scale, UTMALDG = TMA tensor load, UTMAPF = TMA L2 prefetch, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, DMMA = FP64 tensor MMA"
printf "%-110s %8s %8s %8s %7s %6s %7s %6s\n" kernel UTCIMMA UTCOMMA UTMALDG UTMAPF LDTM UTCBAR DMMA
cuobjdump -sass "$LIB" | grep -E "Function :|UTCIMMA|UTCOMMA|UTMALDG|UTMAPF|LDTM|UTCBAR|DMMA" | c++filt | awk '
/Function :/ { if (name != "") out(); name=$0; sub(/.*Function : /, "", name); sub(/\(.*/, "", name); gsub(/^void /, "", name); a=b=c=d=e=f=g=0; next }
/UTCIMMA/ {a++} /UTCOMMA/ {g++} /UTMALDG/ {b++} /UTMAPF/ {c++} /LDTM/ {d++} /UTCBAR/ {e++} /DMMA/ {f++}
function out() { if (a+b+c+d+e+f+g > 0) printf "%-110s %8d %8d %8d %7d %6d %7d %6d\n", substr(name,1,110), a, g, b, c, d, e, f }
END { out() }' | sort
