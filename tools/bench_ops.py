"""Per-op kernels (the reference's ImageOp::run surface, one kernel + one OpBuffer per op) on a C2 frame: time per op
with CUDA events, achieved HBM GB/s against each op's own read+write bytes.  python tools/bench_ops.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import common  # noqa: E402
import imagepipe_b200 as ip  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
W, H = 6000, 4000
stream = torch.cuda.Stream()
ctx = ip.Context(0, stream.cuda_stream)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
frame = ip.synth_cfa_u16(common.SEED, W, 0, H, ctx=ctx)
p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(frame, width=W, height=H, cpp=1), ctx=ctx)
common.fill_ipb_ops(p.ops, common.raw_params())
p.output_size()
g = p.globals
px = W * H
# (name, callable(prev) -> buffer, bytes read + written per pixel)
chain = [("gofloat", lambda b: p.ops.gofloat.run(g), 2 + 4), ("demosaic", lambda b: p.ops.demosaic.run(g, b), 4 + 16),
         ("to_lab", lambda b: p.ops.tolab.run(g, b), 16 + 12), ("basecurve", lambda b: p.ops.basecurve.run(g, b), 12 + 12),
         ("from_lab", lambda b: p.ops.fromlab.run(g, b), 12 + 12), ("gamma", lambda b: p.ops.gamma.run(g, b), 12 + 12)]
buf = None
total = 0.0
with torch.cuda.stream(stream):
    for name, fn, bpp in chain:
        out = fn(buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            out = fn(buf)
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        total += us
        gbs = bpp * px / us / 1e3
        print(f"{name:10s} {us:8.1f} us  {bpp:3d} B/px  {gbs:7.1f} GB/s = {100 * gbs / peak:5.1f}% of {peak:.0f}")
        buf = out
print(f"chain      {total:8.1f} us  ({px / total:.0f} MP/s op by op, allocation included)")
p.set_fused(False)
with torch.cuda.stream(stream):
    p.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        p.run()
    e1.record(stream)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
print(f"Pipeline.run op by op: {us:.1f} us/frame = {px / us:.0f} MP/s")
