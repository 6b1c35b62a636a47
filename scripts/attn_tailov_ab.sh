#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_tailov_ab.log
: > $L
B=build/selftest_attn
timeout 120 $B >> $L 2>&1
for rep in 1 2 3; do for ov in 0 1; do
  echo "== B=16 TAIL_OVERLAP=$ov" >> $L
  DCLIP_ATTN_TAIL_OVERLAP=$ov timeout 60 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done; done
for ov in 0 1; do
  echo "== B=37 TAIL_OVERLAP=$ov" >> $L
  DCLIP_ATTN_TAIL_OVERLAP=$ov timeout 60 $B prof2 37 12 2049 2>&1 | grep -v device >> $L
  echo "== B=8 TAIL_OVERLAP=$ov" >> $L
  DCLIP_ATTN_TAIL_OVERLAP=$ov timeout 60 $B prof2 8 12 2049 2>&1 | grep -v device >> $L
  echo "== 512x512 (16,12,1025) TAIL_OVERLAP=$ov" >> $L
  DCLIP_ATTN_TAIL_OVERLAP=$ov timeout 60 $B prof2 16 12 1025 2>&1 | grep -v device >> $L
done
