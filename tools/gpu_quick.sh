#!/bin/bash
# Quick GPU visit: parity tests, a short bench, and one full ncu capture of the C2 fused kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
for w in ${PROF:-c2}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -f -o gpurun_out/prof_$w \
    python tools/run_frames.py $w 3 > gpurun_out/prof_$w.log 2>&1
done
