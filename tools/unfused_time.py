"""Times the op-by-op path (Pipeline.run with the fused kernels off: one kernel and one OpBuffer per op, to_lab+basecurve
and from_lab+gamma paired) on the C2 frame, device resident: CUDA events around 8 runs after 2 warm-ups."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
import imagepipe_b200 as ip

W, H = 6000, 4000
stream = torch.cuda.Stream()
ctx = ip.Context(0, stream=stream.cuda_stream)
d = ip.synth_cfa_u16(common.SEED, W, 0, H, ctx=ctx)
p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(d, W, H), ctx=ctx)
common.fill_ipb_ops(p.ops, common.raw_params())
p.set_fused(False)
for _ in range(2):
    b = p.run(); del b
ctx.synchronize()
n0 = ctx.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream)
    for _ in range(8):
        b = p.run(); del b
    e1.record(stream)
e1.synchronize()
print(f"op-by-op Pipeline.run, C2 frame: {e0.elapsed_time(e1) / 8 * 1000:.1f} us per frame, {(ctx.launch_count - n0) // 8} kernels")
