#!/bin/bash
# session-3 call B: GPU tests with the folded LayerNorm as default, per-kernel ncu metrics of one forward + predict at B = 16,
# one --set full capture of the attention kernel inside the forward (DRAM traffic for bench.py's roofline.traffic)
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_pytest_d.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_d.log
tail -3 $O/r02_pytest_d.log
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct"
PROF_PREDICT=1 timeout 900 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/fwd_metrics \
    python scripts/prof_forward.py 16 > $O/ncu_fwd.log 2>&1
ncu -i /tmp/fwd_metrics.ncu-rep --page raw --csv > $O/r02_forward_ncu_raw.csv 2>/dev/null
python scripts/ncu_table.py $O/r02_forward_ncu_raw.csv > $O/r02_forward_ncu_table.txt 2>&1
head -60 $O/r02_forward_ncu_table.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_fwd_persistent \
    --launch-skip 3 --launch-count 1 -f -o $O/r02_attn_in_forward python scripts/prof_forward.py 16 > $O/ncu_attn_fwd.log 2>&1
python scripts/ncu_hot.py $O/r02_attn_in_forward.ncu-rep 30 > $O/r02_attn_in_forward_hot.txt 2>&1
head -30 $O/r02_attn_in_forward_hot.txt
