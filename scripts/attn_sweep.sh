#!/bin/bash
# A/B sweep of the flash-attention kernel variants (build/selftest_attn from csrc/selftest_attn.cu) on one B200.
# Usage: gpurun -- 'bash scripts/attn_sweep.sh'   -> gpurun_out/attn_sweep.log
mkdir -p gpurun_out
L=gpurun_out/attn_sweep.log
: > $L
B=build/selftest_attn
echo "== correctness, defaults (DEFER=1 SPEC=0 POLY=0 TOKEN=1 TAIL=4)" >> $L
timeout 300 $B >> $L 2>&1
echo "== correctness, DEFER=1 SPEC=1 POLY=1" >> $L
DCLIP_ATTN_SPEC=1 DCLIP_ATTN_POLY=1 timeout 300 $B >> $L 2>&1
for defer in 0 1; do for spec in 0 1; do for poly in 0 1; do for tok in 1 0; do for tail in 4 0; do
  if [ $tail = 0 ] && { [ $poly != 0 ] || [ $tok != 1 ] || [ $spec != 0 ]; }; then continue; fi
  echo "== DEFER=$defer SPEC=$spec POLY=$poly TOKEN=$tok TAIL=$tail" >> $L
  DCLIP_ATTN_DEFER=$defer DCLIP_ATTN_SPEC=$spec DCLIP_ATTN_POLY=$poly DCLIP_ATTN_TOKEN=$tok DCLIP_ATTN_TAIL_ROWS=$tail timeout 120 $B prof2 16 12 2049 2>&1 | grep -v device >> $L
done; done; done; done; done
echo "== timeline DEFER=1" >> $L
timeout 120 build/selftest_attn_tl timeline 16 12 2049 >> $L 2>&1
echo "== timeline DEFER=1 POLY=1" >> $L
DCLIP_ATTN_POLY=1 timeout 120 build/selftest_attn_tl timeline 16 12 2049 >> $L 2>&1
echo "== timeline DEFER=1 SPEC=1" >> $L
DCLIP_ATTN_SPEC=1 timeout 120 build/selftest_attn_tl timeline 16 12 2049 >> $L 2>&1
echo "== tail-only launches (192 CTAs): CUDA-core tail path vs tensor-core path" >> $L
DCLIP_ATTN_TAIL_ROWS=4 timeout 60 $B qstart 16 12 2049 2048 2>&1 | grep -v device >> $L
DCLIP_ATTN_TAIL_ROWS=0 timeout 60 $B qstart 16 12 2049 2048 2>&1 | grep -v device >> $L
DCLIP_ATTN_TAIL_ROWS=0 timeout 60 $B qstart 16 12 2049 1792 2>&1 | grep -v device >> $L
echo "== L14 shape" >> $L
for defer in 0 1; do DCLIP_ATTN_DEFER=$defer timeout 120 $B prof2 8 16 2629 2>&1 | grep -v device >> $L; done
tail -3 $L
