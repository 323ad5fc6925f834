#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL stripe test, C5 strong-scaling bench at 1..N, C2 replicas at N.
set -u
N=${1:-2}
mkdir -p gpurun_out
python tools/pcie_probe.py 72 > gpurun_out/pcie.log 2>&1
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_sharded.log
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/c5_n1.json 2> gpurun_out/c5_n1.err; echo "c5 n1 rc=$?"
n=2
while [ $n -le $N ]; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --workload c5 --steps 10 --warmup 3 > gpurun_out/c5_n$n.json 2> gpurun_out/c5_n$n.err; echo "c5 n$n rc=$?"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_n$n.json 2> gpurun_out/c2_n$n.err; echo "c2 n$n rc=$?"
  n=$((n * 2))
done
cat gpurun_out/pcie.log gpurun_out/c5_n*.json gpurun_out/c2_n*.json
grep -h -B2 -A12 "Traceback" gpurun_out/*.err | head -60
