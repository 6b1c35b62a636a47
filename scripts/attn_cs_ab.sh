#!/bin/bash
# Column-split attention kernel vs the two-warpgroup persistent kernel: correctness (selftest) and same-box timing.
L=gpurun_out/r02_attn_cs_ab.log
: > $L
echo "## selftest (DCLIP_ATTN_CS=1)" >> $L
DCLIP_ATTN_CS=1 timeout 180 build/selftest_attn >> $L 2>&1; echo "rc=$?" >> $L
for rep in 1 2; do
for v in "DCLIP_ATTN_CS=0" "DCLIP_ATTN_CS=1" "DCLIP_ATTN_CS=1 DCLIP_ATTN_POLY=1" "DCLIP_ATTN_CS=1 DCLIP_ATTN_POLY=2"; do
  echo "## $v prof2 16 12 2049" >> $L
  env $v timeout 60 build/selftest_attn prof2 16 12 2049 2>&1 | tail -1 >> $L
done
done
for v in "DCLIP_ATTN_CS=0" "DCLIP_ATTN_CS=1" "DCLIP_ATTN_CS=1 DCLIP_ATTN_POLY=1"; do
  echo "## $v prof2 8 16 2629 (ViT-L/14)" >> $L
  env $v timeout 60 build/selftest_attn prof2 8 16 2629 2>&1 | tail -1 >> $L
  echo "## $v prof2 37 12 2049" >> $L
  env $v timeout 60 build/selftest_attn prof2 37 12 2049 2>&1 | tail -1 >> $L
done
cat $L
