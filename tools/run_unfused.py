"""Op-by-op Pipeline::run of a C2 frame (one kernel + one OpBuffer per op): the command profiled for the per-op kernels.
  python tools/run_unfused.py [scaled]      scaled: with the 4x down-scale (k_transform_buffer<CFA> instead of k_demosaic_full)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common, imagepipe_b200 as ip
ctx = ip.Context(0)
W, H = 6000, 4000
frame = ip.synth_cfa_u16(common.SEED, W, 0, H, ctx=ctx)
p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(frame, width=W, height=H, cpp=1), ctx=ctx)
common.fill_ipb_ops(p.ops, common.raw_params())
if len(sys.argv) > 1 and sys.argv[1] == "scaled":
    p.globals.settings.maxwidth, p.globals.settings.maxheight = 1500, 1000
p.set_fused(False)
for _ in range(2):
    b = p.run()
ctx.synchronize()
