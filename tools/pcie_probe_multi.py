"""Aggregate host<->device copy ceiling of the box when N GPUs copy at once — the ceiling of bench.py's e2e leg at N GPUs.

  python -m torch.distributed.run --nproc-per-node N tools/pcie_probe_multi.py

Every rank moves the e2e frame's traffic (48 MB host->device, 72 MB device->host, pinned memory, two streams, both
directions at once) in a loop between two barriers; the line printed by rank 0 gives the per-GPU and aggregate GB/s and
the frame rate that traffic allows: ceiling_mpps = N * 24 MP / (time per frame).  Also prints the NUMA node and CPU
affinity of every GPU as the driver reports them."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
IN, OUT, REPS = 48_000_000, 72_000_000, 20
h_in = torch.empty(IN, dtype=torch.uint8).pin_memory()
h_out = torch.empty(OUT, dtype=torch.uint8).pin_memory()
d_in = torch.empty(IN, dtype=torch.uint8, device="cuda")
d_out = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def frame():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for _ in range(3):
    frame()
sync_all()
t0 = time.perf_counter()
for _ in range(REPS):
    frame()
torch.cuda.synchronize()
t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
per_frame = float(t.item()) / REPS
numa = None
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    with open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node") as f:
        numa = int(f.read())
except Exception:
    pass
info = [None] * world
mine = {"gpu": local, "numa_node": numa, "cpus_allowed": len(os.sched_getaffinity(0))}
if world > 1:
    dist.all_gather_object(info, mine)
else:
    info = [mine]
if rank == 0:
    print(json.dumps({"n_gpus": world, "ms_per_frame": per_frame * 1e3, "h2d_gbs_per_gpu": IN / per_frame / 1e9,
                      "d2h_gbs_per_gpu": OUT / per_frame / 1e9, "aggregate_gbs": world * (IN + OUT) / per_frame / 1e9,
                      "ceiling_mpps": world * 24.0 / per_frame, "host_cpus": os.cpu_count(), "gpus": info}))
if world > 1:
    dist.destroy_process_group()
