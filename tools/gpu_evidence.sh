#!/bin/bash
# Evidence visit (1 GPU): sanitizer passes at HEAD, ncu --set full of the per-op demosaic and resampler kernels, the ncu
# launch list of the bench command, per-op kernel durations.
set -u
mkdir -p gpurun_out
bash tools/sanitize.sh > gpurun_out/sanitize_tail.txt 2>&1
tail -14 gpurun_out/sanitize_tail.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_demosaic_full -c 1 -f -o gpurun_out/prof_demosaic_full \
  python tools/run_unfused.py > gpurun_out/prof_demosaic_full.log 2>&1; echo "ncu demosaic rc=$?"
python tools/ncu_summary.py gpurun_out/prof_demosaic_full.ncu-rep > gpurun_out/prof_demosaic_full.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_transform_buffer -c 1 -f -o gpurun_out/prof_transform_buffer \
  python tools/run_unfused.py scaled > gpurun_out/prof_transform_buffer.log 2>&1; echo "ncu transform rc=$?"
python tools/ncu_summary.py gpurun_out/prof_transform_buffer.ncu-rep > gpurun_out/prof_transform_buffer.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --frames-per-step 8 --no-cpu-baseline --no-strong > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_unfused.csv \
  python tools/run_unfused.py > /dev/null 2>&1
head -30 gpurun_out/prof_demosaic_full.txt; head -30 gpurun_out/prof_transform_buffer.txt
