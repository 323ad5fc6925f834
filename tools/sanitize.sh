#!/bin/bash
# compute-sanitizer passes over small parity tests: memcheck (out-of-bounds / misaligned accesses in every kernel the
# tests launch) and racecheck (shared-memory hazards of the tile pipeline of the fused kernels).
set -u
mkdir -p gpurun_out
SEL='full_bayer_f32 or scaled_pipeline or row_stripes or full_other_cfas or tma_and_plain'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py tests/test_gpu_lanczos.py tests/test_gpu_ops.py -m gpu -x -q -k "$SEL or lanczos_matches or demosaic_full or gofloat_raw or tolab" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "full_bayer_f32 or scaled_pipeline" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log
tail -6 gpurun_out/sanitizer_memcheck.log; tail -6 gpurun_out/sanitizer_racecheck.log
