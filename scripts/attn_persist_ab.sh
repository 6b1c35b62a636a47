#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_persist_ab.log
: > $L
B=build/selftest_attn
echo "== selftest, persistent default" >> $L
timeout 120 $B >> $L 2>&1; echo "rc=$?" >> $L
echo "== selftest, persistent POLY=1" >> $L
DCLIP_ATTN_POLY=1 timeout 120 $B 2>&1 | tail -4 >> $L
for rep in 1 2; do for pers in 0 1; do
  echo "== B=16 PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done; done
for cfg in "1 1 0" "1 1 1" "1 0 0"; do set -- $cfg
  echo "== B=16 PERSIST=1 POLY=$1 TOKEN=$2 PEEL=$3" >> $L
  DCLIP_ATTN_POLY=$1 DCLIP_ATTN_TOKEN=$2 DCLIP_ATTN_PEEL=$3 timeout 60 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done
for pers in 0 1; do
  echo "== B=37 PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 37 12 2049 2>&1 | grep -v device >> $L
  echo "== L14 (8,16,2629) PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 8 16 2629 2>&1 | grep -v device >> $L
  echo "== 512x512 (16,12,1025) PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 16 12 1025 2>&1 | grep -v device >> $L
  echo "== B=8 (8,12,2049) PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 8 12 2049 2>&1 | grep -v device >> $L
done
echo "== timeline persistent" >> $L
DCLIP_TL_ROWS=24 DCLIP_TL_CTA=70 timeout 60 build/selftest_attn_tl timeline 16 12 2049 2>&1 | tail -49 >> $L
