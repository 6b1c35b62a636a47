#!/bin/bash
O=gpurun_out
for v in 1 0; do
  DENSECLIP_B200_LN_FOLD=$v timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16 \
      --launch-skip 5 --launch-count 3 -f -o $O/r02_fold${v}_gemms python scripts/prof_forward.py 16 > $O/ncu_fold$v.log 2>&1
  python scripts/ncu_hot.py $O/r02_fold${v}_gemms.ncu-rep 40 > $O/r02_fold${v}_gemms_hot.txt 2>&1
done
grep -A3 "== kernel" $O/r02_fold1_gemms_hot.txt | head -20
