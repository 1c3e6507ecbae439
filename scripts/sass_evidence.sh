#!/bin/bash
# SASS evidence that the product library is Blackwell-native: per kernel family, how many tcgen05 MMA (UTC*MMA),
# TMA load / store / reduce (UTMALDG / UTMASTG / UTMAREDG), TMEM load (LDTM) and legacy mma.sync (HMMA) instructions it has.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LIB=$ROOT/marian-nmt-distributed_b200/lib/libmarian_b200.so
echo "# cuobjdump -sass $(basename $LIB) ($(date -u +%Y-%m-%d)), instruction counts per kernel family"
echo "| kernel family | variants | UTC*MMA | UTMALDG | UTMASTG | UTMAREDG | LDTM | UTCBAR | HMMA (mma.sync) |"
echo "|---|---:|---:|---:|---:|---:|---:|---:|---:|"
cuobjdump -sass $LIB 2>/dev/null | awk '
/Function :/ { fn=$3; fam=fn;
  if (fn ~ /gGemmBf16Persistent/) fam="gGemmBf16Persistent"; else if (fn ~ /gGemmBf16/) fam="gGemmBf16"; else if (fn ~ /gGemmTf32/) fam="gGemmTf32";
  else if (fn ~ /gGemmTcgen05/) fam="gGemmTcgen05 (packed bf16 / bf16x3)"; else if (fn ~ /gAttention/) fam="gAttention*"; else fam="(other kernels)";
  n[fam]++ }
/UTC[A-Z]*MMA/ { mma[fam]++ } /UTMALDG/ { ldg[fam]++ } /UTMASTG/ { stg[fam]++ } /UTMAREDG/ { red[fam]++ } /LDTM/ { ldtm[fam]++ } /UTCBAR/ { bar[fam]++ } /[ \t]HMMA\./ { hmma[fam]++ }
END { for (f in n) printf "| %s | %d | %d | %d | %d | %d | %d | %d | %d |\n", f, n[f], mma[f], ldg[f], stg[f], red[f], ldtm[f], bar[f], hmma[f] }' | sort
