#!/bin/bash
# --set full captures of the remaining kernels of the forward at B = 16 (one launch each): grouped neck conv, LayerNorm (ln_pre),
# upsample, upsample + argmax, score map, patch im2col
O=gpurun_out
# (filters are regular expressions over the DEMANGLED kernel name, template arguments included)
for k in "gemm_bf16_tcgen05_kernel.*int.256.*int.4.*int.68:neckconv" "layernorm_kernel.*int.6:layernorm" "upsample_bilinear_tok_kernel:upsample" "upsample_argmax_strip_kernel:upsample_argmax" "score_map_kernel:score_map" "im2col_patch_vec8_kernel:im2col" "attn_small_kernel.*int.20:attn_small"; do
  re="$(echo "$k" | cut -d: -f1)"; name="$(echo "$k" | cut -d: -f2)"
  PROF_PREDICT=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$re" --launch-count 1 -f -o /tmp/r02_full_$name python scripts/prof_forward.py 16 > $O/ncu_full_$name.log 2>&1
  python scripts/ncu_hot.py /tmp/r02_full_$name.ncu-rep 12 > $O/r02_full_$name.txt 2>&1
  echo "== $name"; grep "gpu__time_duration\|dram__bytes\|tensor_cycles\|dram_throughput" $O/r02_full_$name.txt | head -5
done
