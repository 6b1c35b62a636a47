#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_quant_ab.log
: > $L
B=build/selftest_attn
for peel in 0 1; do for tail in 0 4; do
  echo "== B=37 (24 full waves of regular CTAs) PEEL=$peel TAIL=$tail" >> $L
  DCLIP_ATTN_PEEL=$peel DCLIP_ATTN_TAIL_ROWS=$tail timeout 120 $B prof2 37 12 2049 2>&1 | grep -v device >> $L
done; done
for peel in 0 1; do
  echo "== q_start=1 N=2049 B=37 H=4 (8 full blocks only: 1184 CTAs = 8 waves) PEEL=$peel" >> $L
  DCLIP_ATTN_PEEL=$peel timeout 60 $B qstart 37 4 2049 1 2>&1 | grep -v device >> $L
  echo "== q_start=1 N=2049 B=16 H=12 (1536 CTAs = 10.4 waves) PEEL=$peel" >> $L
  DCLIP_ATTN_PEEL=$peel timeout 60 $B qstart 16 12 2049 1 2>&1 | grep -v device >> $L
done
