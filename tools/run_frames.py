"""Run a few frames of one workload through the C ABI (device-resident in/out): the command profiled under ncu.
  python tools/run_frames.py [c2|c3|c4|c2f32] [nframes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
import imagepipe_b200 as ip  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = ip.Context(0)
if which == "c3":
    w, h, cfa, st = 8256, 5504, common.XTRANS, {}
elif which == "c4":
    w, h, cfa, st = 6000, 4000, "RGGB", {"maxwidth": 1500, "maxheight": 1000}
else:
    w, h, cfa, st = 6000, 4000, "RGGB", {}
nsets = int(os.environ.get("IPB_SETS", "2"))  # rotating input / output sets (8: steady-state DRAM traffic, write-back included)
frames = [ip.synth_cfa_u16(common.SEED + i, w, 0, h, ctx=ctx) for i in range(nsets)]
dsts = [ip.DeviceArray(w * h * 3 * (4 if which == "c2f32" else 1), ctx) for _ in range(nsets if which != "c2f32" else 1)]
for i in range(n):
    dst = dsts[i % len(dsts)]
    p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(frames[i % nsets], width=w, height=h, cpp=1), ctx=ctx)
    common.fill_ipb_ops(p.ops, common.raw_params(cfa=cfa))
    for k, v in st.items():
        setattr(p.globals.settings, k, v)
    if which == "c2f32":
        buf = p.run()
        del buf
    else:
        p.output_8bit(dst=dst)
ctx.synchronize()
print("ran", n, which, "frames;", ctx.launch_count, "launches")
