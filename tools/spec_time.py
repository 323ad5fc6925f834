"""Times the C2 frame (6000x4000 RGGB -> sRGB8, device resident) through the exact fused kernel and the speculative
kernel at both CTA sizes: CUDA events around 16 launches over 8 rotating buffer sets, fix-up fraction, probe."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import common
import imagepipe_b200 as ip

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (6000, 4000)
NSETS, REPS = 8, 4
stream = torch.cuda.Stream()
ctx = ip.Context(0, stream=stream.cuda_stream)
params = common.raw_params()
sets = []
for k in range(NSETS):
    d = ip.synth_cfa_u16(common.SEED + k, W, 0, H, ctx=ctx)
    p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(d, W, H), ctx=ctx)
    common.fill_ipb_ops(p.ops, params)
    sets.append((p, ip.DeviceArray(W * H * 3, ctx)))


def run(label, spec, threads, delta=0.0):
    ctx.set_spec(delta, threads)
    for p, _ in sets:
        p.set_speculative(spec)
    for p, out in sets:                       # warm-up (also builds the tables)
        p.output_8bit(dst=out)
    ctx.synchronize()
    ctx.spec_stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(REPS):
            for p, out in sets:
                p.output_8bit(dst=out)
        e1.record(stream)
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (REPS * NSETS)
    st = ctx.spec_stats()
    fx = st["fixups"] / (REPS * NSETS) / (W * H)
    print(f"{label:34s} {us:8.1f} us/frame  {W*H/us/1e3:7.1f} GP/s  {5*W*H/us/1e3:7.1f} GB/s  fix-ups {100*fx:.3f} %  delta {st['delta']:.3g}")


run("exact fused kernel", False, 512)
run("speculative, 2 x 512 threads", True, 512)
run("speculative, 1 x 1024 threads", True, 1024)
run("speculative 512, delta 1e-5", True, 512, 1e-5)
run("speculative 512, delta 2e-5", True, 512, 2e-5)
run("speculative 512, delta 7.9e-5", True, 512, 7.9e-5)
ctx.set_spec(0.0, 512)
mx, mean, delta = sets[0][0].spec_probe()
print(f"probe: max |cheap-exact| {mx:.3g}  mean {mean:.3g}  certified delta {delta:.3g}  mufu err {ctx.spec_stats()['mufu_err']:.3g}")
run("speculative 512, delta 1e-7 (cheap pass only)", True, 512, 1e-7)
run("speculative 1024, delta 1e-7 (cheap pass only)", True, 1024, 1e-7)
