#!/bin/bash
# GPU visit for the speculative kernel: its tests, a short bench, per-kernel durations at both CTA sizes.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py -m gpu -x -q > gpurun_out/pytest_spec.log 2>&1; echo "pytest spec rc=$?"
tail -25 gpurun_out/pytest_spec.log
timeout 300 python tools/spec_time.py > gpurun_out/spec_time.txt 2>&1; echo "spec_time rc=$?"
cat gpurun_out/spec_time.txt
