"""Times n stripes (or whole frames) of one geometry through k_spec8: one batched launch against n single launches, both
replayed from a CUDA graph (device resident).  python tools/batch_time.py [width rows nframes stripe|whole]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
import imagepipe_b200 as ip
from imagepipe_b200 import _capi
from imagepipe_b200.sharded import DevicePtr

W = int(sys.argv[1]) if len(sys.argv) > 1 else 11648
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1092
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10
stripe = (sys.argv[4] if len(sys.argv) > 4 else "stripe") == "stripe"
H = 8736 if stripe else R
stream = torch.cuda.Stream()
ctx = ip.Context(0, stream=stream.cuda_stream)
r0 = 3276 if stripe else 0
r1 = r0 + R
dummy = ip.DeviceArray(64, ctx)
p = ip.Pipeline.new_from_source(ip.ImageSource(_capi.SRC_RAW_U16, W, H, 1, dummy), ctx=ctx)
common.fill_ipb_ops(p.ops, common.raw_params())
s0, s1 = p.stripe_rows(r0, r1) if stripe else (0, H)
rows = s1 - s0
with torch.cuda.stream(stream):
    src = torch.zeros((N, rows, W), dtype=torch.int16, device="cuda")
    dst = torch.empty((N, R, W, 3), dtype=torch.uint8, device="cuda")
    for k in range(N):
        ip.lib().ipb_synth_cfa_u16(ctx.handle, common.SEED + k, W, s0, rows, src[k].data_ptr())
p.set_stripe_source(ip.ImageSource(_capi.SRC_RAW_U16, W, rows, 1, src.data_ptr()), s0, r0, r1)
nbytes = R * W * 3
out_all = DevicePtr(dst.data_ptr(), dst.numel())


def batched():
    p.output_8bit_batch(N, rows, out_all, nbytes)


def singles():
    for k in range(N):
        p.set_stripe_source(ip.ImageSource(_capi.SRC_RAW_U16, W, rows, 1, src[k].data_ptr()), s0, r0, r1)
        p.output_8bit_stripe(dst=DevicePtr(dst[k].data_ptr(), nbytes), rows=R, width=W)


for name, fn in (("one batched launch", batched), ("single launches", singles)):
    with torch.cuda.stream(stream):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
        fn()
    with torch.cuda.stream(stream):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(10):
            g.replay()
        e1.record(stream)
    e1.synchronize()
    print(f"{W}x{R} x {N} {'stripes' if stripe else 'frames'}, {name:20s}: {e0.elapsed_time(e1) * 100 / N:8.1f} us per frame")
    p.set_stripe_source(ip.ImageSource(_capi.SRC_RAW_U16, W, rows, 1, src.data_ptr()), s0, r0, r1)
