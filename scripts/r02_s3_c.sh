#!/bin/bash
# session-3 call C: full GPU suite, smoke, default bench (graph), reference arm
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_e.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_e.log
tail -4 $O/r02_pytest_e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $O/r02_bench_final_a.json 2> $O/r02_bench_final_a.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$O/r02_bench_final_a.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "enc ms", round(d["encoder"]["ms_per_step"], 3),
      "enc sustained frac", round(d["encoder"]["frac_of_sustained_peak"], 3), "attn", d["roofline"]["ms_per_launch"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"],
      "clk", d["clocks"], "launches", d["gpu_launches"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2>> $O/r02_bench_final_a.err; echo "ref rc=$?"; cut -c1-300 $O/r02_bench_reference.json
