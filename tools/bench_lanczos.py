"""Lanczos-3 extension op on the f32x3 result of a C2 frame (6000x4000 -> 1500x1000): time per call (two launches,
table build on the host included) and the two kernels' share of HBM bandwidth.  python tools/bench_lanczos.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import common  # noqa: E402
import imagepipe_b200 as ip  # noqa: E402

stream = torch.cuda.Stream()
ctx = ip.Context(0, stream.cuda_stream)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
W, H = 6000, 4000
frame = ip.synth_cfa_u16(common.SEED, W, 0, H, ctx=ctx)
p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(frame, width=W, height=H, cpp=1), ctx=ctx)
common.fill_ipb_ops(p.ops, common.raw_params())
buf = p.run()
for nw, nh in ((1500, 1000), (3000, 2000), (750, 500)):
    with torch.cuda.stream(stream):
        for _ in range(2):
            out = ip.lanczos_resize(buf, nw, nh)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(stream)
        for _ in range(reps):
            out = ip.lanczos_resize(buf, nw, nh)
        e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    algo = 4 * 3 * (W * H + 2 * nw * H + nw * nh)   # read source, write + read intermediate, write result
    print(f"lanczos3 {W}x{H}x3 f32 -> {nw}x{nh}: {us:8.1f} us/call  {W * H / us:8.0f} MP/s  "
          f"{algo / us / 1e3:7.1f} GB/s algorithmic = {100 * algo / us / 1e3 / peak:5.1f}% of {peak:.0f} GB/s")
