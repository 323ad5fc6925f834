#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), all-config timings, ncu launch list of the bench command,
# full ncu capture of the fused kernels for C2 / C3 / C4.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/bench_configs.py > gpurun_out/bench_configs.txt 2>&1
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for w in c2 c3 c4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -f -o gpurun_out/prof_$w \
    python tools/run_frames.py $w 3 > gpurun_out/prof_$w.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_configs.txt
