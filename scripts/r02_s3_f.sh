#!/bin/bash
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_train_tail.py -m gpu -q 2>&1 | tail -2
PROF_TRAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct \
    --clock-control none --profile-from-start off -f -o /tmp/train_metrics python scripts/prof_train_step.py 16 > $O/ncu_train.log 2>&1
ncu -i /tmp/train_metrics.ncu-rep --page raw --csv > $O/r02_train_ncu_raw.csv 2>/dev/null
python scripts/ncu_table.py $O/r02_train_ncu_raw.csv > $O/r02_train_step_ncu_table.txt 2>&1
head -24 $O/r02_train_step_ncu_table.txt
