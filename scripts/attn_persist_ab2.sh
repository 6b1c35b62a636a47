#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_persist_ab2.log
: > $L
B=build/selftest_attn
timeout 120 $B 2>&1 | tail -4 >> $L
for rep in 1 2; do for pers in 0 1; do
  echo "== B=16 PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done; done
for pers in 0 1; do
  echo "== B=37 PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 37 12 2049 2>&1 | grep -v device >> $L
  echo "== L14 (8,16,2629) PERSIST=$pers" >> $L
  DCLIP_ATTN_PERSIST=$pers timeout 60 $B prof2 8 16 2629 2>&1 | grep -v device >> $L
done
echo "== timeline persistent (rows = tiles 14.. of item 0; row 31 = output pass of item 1: wait_start=begin s_full_got=z0 done ld_done=o_done got max_done=tile stored exps_done=barrier passed arrived=store issued)" >> $L
DCLIP_TL_ROWS=32 DCLIP_TL_CTA=70 timeout 60 build/selftest_attn_tl timeline 16 12 2049 2>&1 | tail -64 >> $L
