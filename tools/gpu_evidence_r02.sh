#!/bin/bash
# Round-2 evidence visit (1 GPU) at HEAD: all GPU tests, the default bench line and the reference arm, the C3 / C4 lines,
# ncu --set full of the three speculative kernels (+ traffic passes), the launch list of the bench command, sanitizers,
# per-op kernel durations, the device-resident timings of every configuration.
set -u
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for w in c3 c4; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; done
python tools/bench_configs.py > gpurun_out/bench_configs.txt 2>&1
python tools/unfused_time.py > gpurun_out/unfused_time.txt 2>&1
python tools/spec_time.py > gpurun_out/spec_time.txt 2>&1
for w in c2 c3 c4; do bash tools/gpu_prof_spec.sh $w > gpurun_out/prof_$w.out 2>&1; python tools/ncu_stalls.py gpurun_out/prof_spec_$w.ncu-rep 2 > gpurun_out/prof_spec_${w}_stalls.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --frames-per-step 8 --no-cpu-baseline --no-strong > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_unfused.csv \
  python tools/unfused_time.py > /dev/null 2>&1
bash tools/sanitize.sh > gpurun_out/sanitize_tail.txt 2>&1; tail -8 gpurun_out/sanitize_tail.txt
cat gpurun_out/bench.json | tail -1 | cut -c1-600; cat gpurun_out/bench_ref.json | tail -1 | cut -c1-400
cat gpurun_out/bench_configs.txt gpurun_out/unfused_time.txt
