#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md: tcgen05.mma ->
# UTC*MMA, cp.async.bulk.tensor -> UTMALDG / UTMASTG, cp.async.bulk -> UBLKCP, tcgen05.ld/st -> LDTM / STTM).
# Runs anywhere cuobjdump is installed (no GPU needed):  bash tools/sass_mnemonics.sh > profiles/<round>_sass_mnemonics.txt
LIB=${1:-vit_exp_b200/libctk.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ {fn=$3; seen[fn]=1}
/UTC[A-Z]*MMA/ {mma[fn]++}
/UTMALDG/ {ldg[fn]++}
/UTMASTG/ {stg[fn]++}
/UBLKCP/ {blk[fn]++}
/LDTM/ {tld[fn]++}
/STTM/ {tst[fn]++}
/HMMA/ && !/UTC/ {hm[fn]++}
/MUFU.EX2/ {ex2[fn]++}
END {for (f in seen) if (mma[f] + ldg[f] + stg[f] + blk[f] + hm[f] > 0)
  printf "%s | UTCxMMA %d | UTMALDG %d | UTMASTG %d | UBLKCP %d | LDTM %d | STTM %d | HMMA %d | MUFU.EX2 %d\n", f, mma[f], ldg[f], stg[f], blk[f], tld[f], tst[f], hm[f], ex2[f]}' \
 | c++filt | sed -e 's/(anonymous namespace):://g' -e 's/(CUtensorMap_st.*) |/(...) |/' -e 's/(__nv_bfloat16.*) |/(...) |/' -e 's/(float const\*.*) |/(...) |/' -e 's/^void //' | sort
