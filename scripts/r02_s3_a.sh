#!/bin/bash
# session-3 call A: full GPU tests, LayerNorm-fold parity + same-box A/B, GEMM captures with the hoisted epilogue loads
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_pytest_c.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_c.log
tail -3 $O/r02_pytest_c.log
DENSECLIP_B200_LN_FOLD=1 timeout 600 python -m pytest tests/test_gpu_parity_baseline.py tests/test_gpu_denseclip.py tests/test_gpu_fullsize.py -m gpu -x -q -s > $O/r02_pytest_fold.log 2>&1; echo "pytest fold rc=$?" >> $O/r02_pytest_fold.log
grep "PARITY\|passed\|failed\|rc=" $O/r02_pytest_fold.log | cut -c1-400
for rep in 1 2; do
for v in 0 1; do
  DENSECLIP_B200_LN_FOLD=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_fold${v}_s3_$rep.json 2> $O/r02_bench_fold${v}_s3.err
  python - <<PY
import json
d = json.loads([l for l in open("$O/r02_bench_fold${v}_s3_$rep.json") if l.startswith("{")][-1])
print("fold=$v rep=$rep value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "encoder ms", round(d["encoder"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
done
done
for v in 1 0; do
  DENSECLIP_B200_LN_FOLD=$v timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16 \
      --launch-skip 5 --launch-count 4 -f -o $O/r02_fold${v}b_gemms python scripts/prof_forward.py 16 > $O/ncu_fold$v.log 2>&1
  python scripts/ncu_hot.py $O/r02_fold${v}b_gemms.ncu-rep 25 > $O/r02_fold${v}b_gemms_hot.txt 2>&1
  grep -A4 "== kernel" $O/r02_fold${v}b_gemms_hot.txt | grep "kernel\|time_duration"
done
