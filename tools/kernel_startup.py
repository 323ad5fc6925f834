"""Fused full-resolution kernel time against frame height (device-resident in/out): the intercept of the fit is the
per-launch start-up cost (tables into shared memory, first tile).  python tools/kernel_startup.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import common  # noqa: E402
import imagepipe_b200 as ip  # noqa: E402

W = 6000
stream = torch.cuda.Stream()
ctx = ip.Context(0, stream.cuda_stream)
res = []
for h in (32, 96, 320, 1000, 4000, 8000):
    frames = [ip.synth_cfa_u16(common.SEED + i, W, 0, h, ctx=ctx) for i in range(4)]
    outs = [ip.DeviceArray(W * h * 3, ctx) for _ in range(4)]
    pipes = []
    for i in range(4):
        p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(frames[i], width=W, height=h, cpp=1), ctx=ctx)
        common.fill_ipb_ops(p.ops, common.raw_params())
        pipes.append(p)
    reps = 40
    with torch.cuda.stream(stream):
        for i in range(8):
            pipes[i % 4].output_8bit(dst=outs[i % 4])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps):
            pipes[i % 4].output_8bit(dst=outs[i % 4])
        e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    tiles = -(-W // 128) * -(-h // 32)
    res.append((h, tiles, us))
    print(f"h={h:5d} tiles={tiles:6d} ({tiles / 148:6.2f} per CTA)  {us:8.1f} us/launch  {W * h / us:9.0f} MP/s")
a, b = np.polyfit([r[1] / 148 for r in res[2:]], [r[2] for r in res[2:]], 1)
print(f"fit over the larger frames: {a:.2f} us per tile-round + {b:.1f} us start-up")
