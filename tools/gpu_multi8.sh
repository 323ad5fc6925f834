#!/bin/bash
# 8-GPU visit: aggregate PCIe ceiling, the default bench line (C2 replicas + C5 strong leg), BASELINE config 4 as written
# (256 frames per step round-robin over 8 GPUs, 4x down-scale).
set -u
N=${1:-8}
mkdir -p gpurun_out
export IPB_BENCH_WATCHDOG=200
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610 \
  tools/pcie_probe_multi.py > gpurun_out/pcie_n$N.json 2> gpurun_out/pcie_n$N.err; echo "pcie rc=$?"
timeout 60 python tools/pcie_probe_multi.py > gpurun_out/pcie_n1.json 2> gpurun_out/pcie_n1.err; echo "pcie1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29620 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29630 \
  bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "c4 rc=$?"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
cat gpurun_out/pcie_n$N.json gpurun_out/pcie_n1.json gpurun_out/bench_n$N.json gpurun_out/bench_c4_n$N.json
grep -h -A12 "Timeout\|Traceback" gpurun_out/*_n$N.err | head -40
