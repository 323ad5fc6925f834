import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import common, imagepipe_b200 as ip
ctx = ip.Context(0)
W, H = 6000, 4000
frame = ip.synth_cfa_u16(common.SEED, W, 0, H, ctx=ctx)
p = ip.Pipeline.new_from_source(ip.ImageSource.Raw(frame, width=W, height=H, cpp=1), ctx=ctx)
common.fill_ipb_ops(p.ops, common.raw_params())
p.set_fused(False)
for _ in range(2):
    b = p.run()
ctx.synchronize()
