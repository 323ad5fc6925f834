"""Host<->device copy bandwidth of the box with pinned memory (the ceiling of bench.py's e2e leg):
H2D alone, D2H alone, both directions at once on two streams.  python tools/pcie_probe.py [MB]"""
import sys

import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 72
n = mb << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    cur.wait_stream(s1)
    cur.wait_stream(s2)


for name, fn, bytes_ in (("H2D", h2d, n), ("D2H", d2h, n), ("H2D+D2H concurrent", both, 2 * n)):
    ms = timed(fn)
    print(f"{name:22s} {mb} MB: {ms:.3f} ms  {bytes_ / ms / 1e6:.1f} GB/s")


def chunked(nchunks):
    def fn():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        c = n // nchunks
        for i in range(nchunks):
            with torch.cuda.stream(s1):
                d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
    return fn


for nc in (4, 16, 64):
    ms = timed(chunked(nc))
    print(f"concurrent, {nc:3d} chunks each way: {ms:.3f} ms  {2 * n / ms / 1e6:.1f} GB/s")
