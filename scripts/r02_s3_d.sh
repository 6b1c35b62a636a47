#!/bin/bash
# session-3 call D: new test, bench with the training step, launch list of the bench command, per-kernel table of one training step
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_train_tail.py -m gpu -q -k "production_shape" 2>&1 | tail -3
timeout 600 python bench.py --train-step --no-cpu-baseline > $O/r02_bench_train.json 2> $O/r02_bench_train.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$O/r02_bench_train.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "enc ms", round(d["encoder"]["ms_per_step"], 3), "train", d.get("train_step"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph > $O/ncu_bench.log 2>&1
python scripts/ncu_launches.py $O/r02_bench_launches.csv > $O/r02_bench_launches.txt 2>&1
head -30 $O/r02_bench_launches.txt
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct"
PROF_TRAIN=1 timeout 900 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/train_metrics \
    python scripts/prof_train_step.py 16 > $O/ncu_train.log 2>&1
ncu -i /tmp/train_metrics.ncu-rep --page raw --csv > $O/r02_train_ncu_raw.csv 2>/dev/null
python scripts/ncu_table.py $O/r02_train_ncu_raw.csv > $O/r02_train_step_ncu_table.txt 2>&1
head -40 $O/r02_train_step_ncu_table.txt
