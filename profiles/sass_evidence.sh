#!/bin/bash
# Which tensor-core / TMEM / bulk-copy SASS mnemonics each kernel of libsoccer2d.so holds (no GPU needed):
#   bash profiles/sass_evidence.sh > profiles/r2_sass_tcgen05.txt
# UTCHMMA = tcgen05.mma (kind::tf32 / f16), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
# HMMA = mma.sync (the actor-critic head and the precision-2 fallback).
SO=${1:-gym-soccer-2d-env_b200/soccer2d_b200/libsoccer2d.so}
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; next }
  { for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|LDTM|STTM|UTCBAR|UTCATOMSWS|HMMA|UBLKCP|UTMALDG|SYNCS)/) { split($i,a,"."); n[fn" "a[1]]++ } }
  END { for (k in n) print n[k], k }' | sort -k2,2 -k3,3 | c++filt | awk '{c=$1; $1=""; printf "%6d %s\n", c, $0}'
