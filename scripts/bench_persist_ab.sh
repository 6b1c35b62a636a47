#!/bin/bash
# same-box A/B of the persistent attention kernel inside the full forward
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for f in nopersist persist nopersist2 persist2; do
  case $f in nopersist*) export DCLIP_ATTN_PERSIST=0;; *) export DCLIP_ATTN_PERSIST=1;; esac
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$f.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms", d["clocks"]["sm_mhz"], "MHz attn", round(d["roofline"]["ms_per_launch"],4), "enc", round(d["encoder"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
