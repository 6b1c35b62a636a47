#!/bin/bash
# session-3 call E: final-code bench lines (default, ViT-L/14, fp32 path, 1-GPU batch sweep, training step)
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --train-step > $O/r02_bench_final_b.json 2> $O/r02_bench_final_b.err; echo "bench rc=$?"
timeout 600 python bench.py --model vit_l14 --batch 8 --no-cpu-baseline > $O/r02_bench_l14_final.json 2>> $O/r02_bench_final_b.err; echo "l14 rc=$?"
timeout 600 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_fp32_final.json 2>> $O/r02_bench_final_b.err; echo "fp32 rc=$?"
timeout 900 python bench.py --sweep 8,32,64,128 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_sweep_1gpu_final.jsonl 2>> $O/r02_bench_final_b.err; echo "sweep rc=$?"
python - <<PY
import json, glob
for f in ["$O/r02_bench_final_b.json", "$O/r02_bench_l14_final.json", "$O/r02_bench_fp32_final.json", "$O/r02_sweep_1gpu_final.jsonl"]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], "B", d["config"]["global_batch"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
                  "enc", round(d["encoder"]["ms_per_step"], 3), round(d["encoder"]["frac_of_sustained_peak"], 3), "attn", round(d["roofline"]["ms_per_launch"], 4), round(d["roofline"]["frac"], 3),
                  "clk", d["clocks"]["sm_mhz"], "train", (d.get("train_step") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
