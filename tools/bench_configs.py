"""Device-resident timing of the BASELINE configurations other than bench.py's headline line (parity-test cases, not
bench lines): C2 with f32x3 output (Pipeline.run), C3 X-Trans full resolution, C4 4x down-scaled output.
CUDA events on the launching stream, rotating buffer sets larger than L2.   python tools/bench_configs.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import common  # noqa: E402
import imagepipe_b200 as ip  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
stream = torch.cuda.Stream()
ctx = ip.Context(0, stream.cuda_stream)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0

CASES = [
    # name, width, height, cfa, settings, output, algorithmic bytes per input pixel
    ("C2 6000x4000 RGGB -> u8x3", 6000, 4000, "RGGB", {}, "u8", 5.0),
    ("C2 6000x4000 RGGB -> f32x3 (Pipeline.run)", 6000, 4000, "RGGB", {}, "f32", 14.0),
    ("C3 8256x5504 X-Trans -> u8x3", 8256, 5504, common.XTRANS, {}, "u8", 5.0),
    ("C4 6000x4000 RGGB -> 1500x1000 u8x3", 6000, 4000, "RGGB", {"maxwidth": 1500, "maxheight": 1000}, "u8", 2.1875),
    ("C5 11648x8736 RGGB -> u8x3 (one GPU)", 11648, 8736, "RGGB", {}, "u8", 5.0),
    # the same kernel on a smooth, natural-looking frame (common.smooth_cfa): nearly every XYZ ratio stays inside the
    # table, so the out-of-table queue and the cube root (23 % of the instructions on the white-noise frames) idle
    ("C2 smooth frame (not white noise) -> u8x3", 6000, 4000, "RGGB", {"smooth": 1}, "u8", 5.0),
]
for name, w, h, cfa, st, out, bpp in CASES:
    nsets = max(2, int(400e6 // (w * h * (2 + (12 if out == "f32" else 3)))) + 1)
    smooth = st.pop("smooth", 0) if isinstance(st, dict) else 0
    if smooth:
        frames = [ip.DeviceArray.from_numpy(common.smooth_cfa(w, h, seed=i + 1), ctx) for i in range(nsets)]
    else:
        frames = [ip.synth_cfa_u16(common.SEED + i, w, 0, h, ctx=ctx) for i in range(nsets)]
    pipes, dsts = [], []
    for i in range(nsets):
        p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(frames[i], width=w, height=h, cpp=1), ctx=ctx)
        common.fill_ipb_ops(p.ops, common.raw_params(cfa=cfa))
        for k, v in st.items():
            setattr(p.globals.settings, k, v)
        pipes.append(p)
        ow, oh = p.output_size()
        dsts.append(ip.DeviceArray(ow * oh * 3, ctx) if out == "u8" else None)

    def one(i):
        if out == "u8":
            pipes[i % nsets].output_8bit(dst=dsts[i % nsets])
        else:
            pipes[i % nsets].run()

    with torch.cuda.stream(stream):
        for i in range(4):
            one(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps):
            one(i)
        e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    gbs = bpp * w * h / us / 1e3
    print(f"{name:46s} {us:8.1f} us/frame  {w * h / us:9.0f} MP/s  {gbs:7.1f} GB/s algorithmic = {100 * gbs / peak:5.2f}% of {peak:.0f} GB/s")
    del pipes, dsts, frames
