#!/bin/bash
# Round-2 ncu evidence, cheapest first, every step under its own timeout.
# usage: gpurun --timeout 1200 -- 'bash scripts/r02_ncu_all.sh'
O=gpurun_out
mkdir -p $O
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct"
# (1) LayerNorm-fold variant vs plain: layer 1's QKV / out-proj / c_fc GEMMs with source correlation
for v in 1 0; do
  DENSECLIP_B200_LN_FOLD=$v timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16 \
      --launch-skip 5 --launch-count 3 -f -o $O/r02_fold${v}_gemms python scripts/prof_forward.py 16 > $O/ncu_fold$v.log 2>&1
  python scripts/ncu_hot.py $O/r02_fold${v}_gemms.ncu-rep 40 > $O/r02_fold${v}_gemms_hot.txt 2>&1
done
# (2) every launch of one eager forward + predict at the bench shape (B = 16), selected metrics only (few replay passes)
PROF_PREDICT=1 timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/fwd_metrics \
    python scripts/prof_forward.py 16 > $O/ncu_fwd.log 2>&1
ncu -i /tmp/fwd_metrics.ncu-rep --page raw --csv > $O/r02_forward_ncu_raw.csv 2>/dev/null
python scripts/ncu_table.py $O/r02_forward_ncu_raw.csv > $O/r02_forward_ncu_table.txt 2>&1
head -40 $O/r02_forward_ncu_table.txt
