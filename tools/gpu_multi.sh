#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL stripe test through the C ABI, the default bench line (C2 replicas + C5 strong
# leg with parity) at N, and the PCIe probe on all GPUs at once.
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_sharded.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
cat gpurun_out/bench_n$N.json
grep -h -B2 -A12 "Traceback" gpurun_out/*.err | head -60
