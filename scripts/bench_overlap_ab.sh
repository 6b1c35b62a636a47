#!/bin/bash
# same-box A/B: score-map / ContextDecoder branch overlapped with the neck / heads branch
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for f in off on off2 on2; do
  case $f in off*) export DCLIP_OVERLAP_TAIL=0;; *) export DCLIP_OVERLAP_TAIL=1;; esac
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ov_$f.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ov_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms", d["clocks"]["sm_mhz"], "MHz attn", round(d["roofline"]["ms_per_launch"],4), "enc", round(d["encoder"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
