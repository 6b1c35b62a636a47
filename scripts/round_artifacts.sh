#!/bin/bash
# Regenerates the per-round GPU artefacts on one B200 (run through gpurun); outputs land in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash scripts/round_artifacts.sh r01'
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${R}_pytest_gpu.log
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
: > $O/${R}_batch_sweep.jsonl
for b in 8 32 64 128; do python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline >> $O/${R}_batch_sweep.jsonl 2>> $O/${R}_bench.err; done
# launch list of the bench command (serialised, cold-cache: shares only), after the plain run above exited
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${R}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph > $O/ncu_bench.log 2>&1
python scripts/ncu_launches.py $O/${R}_bench_launches.csv > $O/${R}_bench_launches.txt 2>&1
# one full capture of the dominant kernel (flash attention at the bench shape) from the stand-alone selftest binary
if [ -x build/selftest_attn ]; then
  build/selftest_attn prof2 16 12 2049 > $O/${R}_selftest_attn_prof.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:attn_fwd --launch-skip 4 --launch-count 1 -f -o $O/${R}_attn_full \
      build/selftest_attn prof2 16 12 2049 > $O/ncu_attn.log 2>&1
  python scripts/ncu_hot.py $O/${R}_attn_full.ncu-rep 30 > $O/${R}_attn_ncu.txt 2>&1
fi
tail -3 $O/${R}_pytest_gpu.log; cut -c1-200 $O/${R}_bench.json; cat $O/${R}_batch_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['global_batch'], round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms', 'e2e', round(d['e2e']['value'], 1))"
