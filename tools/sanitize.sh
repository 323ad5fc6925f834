#!/bin/bash
# compute-sanitizer passes over small parity tests: memcheck (out-of-bounds / misaligned accesses in every kernel the
# tests launch) and racecheck (shared-memory hazards of the tile pipelines: mbarrier / TMA staging, conversion
# counters, the per-tile fix-up queue and the named barrier of the speculative kernels, the warp queues of the scaled
# one; a border-tile case and a cube-root-table case are included).
set -u
mkdir -p gpurun_out
SEL='full_bayer_f32 or scaled_pipeline or row_stripes or full_other_cfas or tma_and_plain or lab_transfer_above_one'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py tests/test_gpu_lanczos.py tests/test_gpu_ops.py tests/test_gpu_spec.py tests/test_gpu_cache.py -m gpu -x -q -k "$SEL or lanczos_matches or demosaic_full or gofloat_raw or tolab or all_phases or queue_flush or frame_kinds or generic_patterns or generic_pattern_crops or scaled_against_oracle or scaled_dark or paired_ops or batch_equals or batch_of_stripes" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py tests/test_gpu_spec.py -m gpu -x -q -k "full_bayer_f32 or scaled_pipeline or full_other_cfas or (all_phases and (130 or 640)) or queue_flush or (generic_patterns and 640 and xtrans) or (generic_pattern_crops and crops0) or (scaled_against_oracle and 640 and RGGB) or scaled_dark or (batch_equals and bayer) or batch_of_stripes" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log
tail -6 gpurun_out/sanitizer_memcheck.log; tail -6 gpurun_out/sanitizer_racecheck.log
