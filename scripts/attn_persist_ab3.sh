#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_persist_ab3.log
: > $L
B=build/selftest_attn
timeout 120 $B 2>&1 | tail -3 >> $L
for rep in 1 2 3; do for peel in 1 2; do for pers in 0 1; do
  echo "== B=16 PERSIST=$pers PEEL=$peel" >> $L
  DCLIP_ATTN_PEEL=$peel DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done; done; done
for pers in 0 1; do
  echo "== L14 (8,16,2629) PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 8 16 2629 2>&1 | grep -v device >> $L
done
