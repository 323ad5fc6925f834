"""Does a cross-stream event dependency chain (H2D_i -> compute_i -> D2H_i) serialise the two copy directions?
Variants: plain chunked copies on two streams; + an event record after every H2D; + the dependency chain; each
timed back to back with CUDA events and with a host synchronize per frame."""
import time

import torch

n = 72 << 20
nin = 48 << 20
h_in = torch.empty(nin, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(nin, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s0, s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def run(nchunks, mode):
    ci, co = nin // nchunks, n // nchunks
    for i in range(nchunks):
        with torch.cuda.stream(s1):
            d_a[i * ci:(i + 1) * ci].copy_(h_in[i * ci:(i + 1) * ci], non_blocking=True)
            if mode >= 1:
                e = torch.cuda.Event()
                e.record(s1)
        if mode >= 2:
            s0.wait_event(e)
            with torch.cuda.stream(s0):
                if mode >= 3:
                    d_b[i * co:(i + 1) * co].fill_(i)
                f = torch.cuda.Event()
                f.record(s0)
            s2.wait_event(f)
        with torch.cuda.stream(s2):
            h_out[i * co:(i + 1) * co].copy_(d_b[i * co:(i + 1) * co], non_blocking=True)


names = {0: "plain", 1: "+event after H2D", 2: "+dependency chain", 3: "+kernel in chain"}
for nchunks in (1, 7, 14):
    for mode in (0, 1, 2, 3):
        for _ in range(2):
            run(nchunks, mode)
        torch.cuda.synchronize()
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            run(nchunks, mode)
            torch.cuda.synchronize()
        ms_sync = (time.perf_counter() - t0) / reps * 1e3
        t0 = time.perf_counter()
        for _ in range(reps):
            run(nchunks, mode)
        torch.cuda.synchronize()
        ms_b2b = (time.perf_counter() - t0) / reps * 1e3
        print(f"chunks={nchunks:2d} {names[mode]:20s}: {ms_sync:.3f} ms/frame with a sync per frame, {ms_b2b:.3f} back to back")
