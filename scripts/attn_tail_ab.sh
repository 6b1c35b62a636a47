#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_tail_ab.log
: > $L
B=build/selftest_attn
DCLIP_ATTN_DEFER=0 DCLIP_ATTN_TAIL_ROWS=4 timeout 300 $B >> $L 2>&1
for rep in 1 2; do for tail in 0 4; do
  echo "== TAIL=$tail" >> $L
  DCLIP_ATTN_DEFER=0 DCLIP_ATTN_TAIL_ROWS=$tail timeout 120 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done; done
DCLIP_ATTN_DEFER=0 DCLIP_ATTN_TAIL_ROWS=4 timeout 60 $B qstart 16 12 2049 2048 2>&1 | grep -v device >> $L
DCLIP_ATTN_DEFER=0 DCLIP_ATTN_TAIL_ROWS=4 timeout 60 $B qstart 2 12 2049 2048 2>&1 | grep -v device >> $L
DCLIP_ATTN_DEFER=0 DCLIP_ATTN_TAIL_ROWS=0 timeout 60 $B qstart 2 12 2049 2048 2>&1 | grep -v device >> $L
