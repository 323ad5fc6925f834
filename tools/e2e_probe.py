"""Time Pipeline.output_8bit with pinned host source and destination (bench.py's e2e leg): T host threads, each with
its own context (private stream), pipeline and pinned buffers, calling the synchronous API in a loop.
  python tools/e2e_probe.py [frames-per-thread] [threads] [band-MB, 0 = whole-frame copies]"""
import ctypes as C
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
import imagepipe_b200 as ip  # noqa: E402

W, H = 6000, 4000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
BAND = int(sys.argv[3]) if len(sys.argv) > 3 else 16


def pinned(nbytes, dtype, shape):
    p = C.c_void_p()
    assert ip.lib().ipb_host_alloc(nbytes, C.byref(p)) == 0
    return np.frombuffer((C.c_uint8 * nbytes).from_address(p.value), dtype=dtype).reshape(shape)


frame = common.synth_cfa(W, H)
workers = []
for t in range(T):
    ctx = ip.Context(0)
    hin, hout = pinned(W * H * 2, np.uint16, (H, W)), pinned(W * H * 3, np.uint8, (H, W, 3))
    hin[:] = frame
    p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(hin), ctx=ctx)
    common.fill_ipb_ops(p.ops, common.raw_params())
    p.set_band_mb(BAND)
    for _ in range(3):
        p.output_8bit(dst=hout)
    workers.append((ctx, p, hout))
start = threading.Barrier(T + 1)


def loop(p, hout):
    start.wait()
    for _ in range(n):
        p.output_8bit(dst=hout)


ths = [threading.Thread(target=loop, args=(w[1], w[2])) for w in workers]
for th in ths:
    th.start()
start.wait()
t0 = time.perf_counter()
for th in ths:
    th.join()
dt = (time.perf_counter() - t0) / (n * T)
print(f"band_mb={BAND} threads={T}: {dt * 1e3:.3f} ms/frame, {W * H / 1e6 / dt:.0f} MP/s")
