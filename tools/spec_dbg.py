"""Timing experiment: certified delta, 2 x 512 threads, with IPB_SPEC_DBG set by the caller."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, common
import imagepipe_b200 as ip
W, H, NSETS, REPS = 6000, 4000, 8, 4
stream = torch.cuda.Stream()
ctx = ip.Context(0, stream=stream.cuda_stream)
sets = []
for k in range(NSETS):
    d = ip.synth_cfa_u16(common.SEED + k, W, 0, H, ctx=ctx)
    p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(d, W, H), ctx=ctx)
    common.fill_ipb_ops(p.ops, common.raw_params())
    sets.append((p, ip.DeviceArray(W * H * 3, ctx)))
for threads in (512, 1024):
    ctx.set_spec(0.0, threads)
    for p, out in sets: p.output_8bit(dst=out)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(REPS):
            for p, out in sets: p.output_8bit(dst=out)
        e1.record(stream)
    e1.synchronize()
    print(f"IPB_SPEC_DBG={os.environ.get('IPB_SPEC_DBG','0')} threads {threads}: {e0.elapsed_time(e1)*1000/(REPS*NSETS):.1f} us/frame")
