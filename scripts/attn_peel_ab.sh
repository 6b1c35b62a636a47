#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_peel_ab.log
: > $L
B=build/selftest_attn
timeout 300 $B >> $L 2>&1
DCLIP_ATTN_PEEL=0 timeout 300 $B 2>&1 | tail -3 >> $L
for rep in 1 2; do for peel in 0 1; do
  echo "== PEEL=$peel" >> $L
  DCLIP_ATTN_PEEL=$peel timeout 120 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done; done
echo "== PEEL=1 POLY=1 TOKEN=0" >> $L
DCLIP_ATTN_POLY=1 DCLIP_ATTN_TOKEN=0 timeout 120 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
echo "== PEEL=1 POLY=1 TOKEN=1" >> $L
DCLIP_ATTN_POLY=1 DCLIP_ATTN_TOKEN=1 timeout 120 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
echo "== 512x512 shape (N=1025) PEEL=0/1" >> $L
for peel in 0 1; do DCLIP_ATTN_PEEL=$peel timeout 120 $B prof2 16 12 1025 2>&1 | grep -v device >> $L; done
echo "== timeline PEEL=1" >> $L
timeout 120 build/selftest_attn_tl timeline 16 12 2049 2>&1 | tail -42 | awk 'NR<4 || NR>30' >> $L
echo "== timeline PEEL=0" >> $L
DCLIP_ATTN_PEEL=0 timeout 120 build/selftest_attn_tl timeline 16 12 2049 2>&1 | tail -8 >> $L
