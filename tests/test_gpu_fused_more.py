"""More GPU parity cases of the fused kernels: base curves of every shape the spline table handles, every window
width of the scaled kernel, four-colour patterns, level mappings that need IEEE division, and the two largest
BASELINE frames (C3 X-Trans 8256x5504, C5 11648x8736) whole against the oracle.  Bit-exact."""
import numpy as np
import pytest

import common
from common import assert_bit_exact

pytestmark = pytest.mark.gpu


def check(ip, orc, ctx, data, params, st=None, fused_expected=True, what=""):
    po = orc.make_pipeline(data, "raw", params, st)
    pg = common.make_ipb_pipeline(ip, data, "raw", params, st, ctx=ctx)
    n0 = ctx.launch_count
    got = pg.run().to_numpy()
    launches = ctx.launch_count - n0
    assert_bit_exact(got, orc.pipeline_run(po), f"{what} f32")
    if fused_expected is not None:
        assert (launches == 1) == fused_expected, f"{what}: {launches} launches"
    po = orc.make_pipeline(data, "raw", params, st)
    pg = common.make_ipb_pipeline(ip, data, "raw", params, st, ctx=ctx)
    assert_bit_exact(pg.output_8bit().to_numpy(), orc.pipeline_output_8bit(po), f"{what} u8")


CURVES = [
    ("passthrough", (), 0.0, True),                                   # no points, no exposure: curves.rs:34-36
    ("exposure only", (), 0.7, True),                                 # end points only, y scaled
    ("default", ((0.5, 0.6),), 0.0, True),
    ("two knots", ((0.3, 0.2), (0.7, 0.9)), 0.0, True),
    ("four knots", ((0.1, 0.05), (0.3, 0.35), (0.6, 0.55), (0.9, 0.95)), -0.3, True),
    ("own end points", ((0.0, 0.1), (0.4, 0.5), (1.0, 0.9)), 0.0, True),   # no (0,0)/(1,1) added: curves.rs:69-75
    ("non-monotone", ((0.2, 0.6), (0.5, 0.3), (0.8, 0.7)), 0.0, True),     # slopes change sign: c1 = 0 branch
    ("unsorted knots", ((0.6, 0.5), (0.3, 0.4)), 0.0, False),              # not fusable: op-by-op binary search
]


@pytest.mark.parametrize("name,points,exposure,fused", CURVES, ids=[c[0] for c in CURVES])
def test_basecurves_through_the_fused_kernel(ip, orc, ctx, name, points, exposure, fused):
    data = common.synth_cfa(403, 131, seed=101)
    check(ip, orc, ctx, data, common.raw_params(points=points, exposure=exposure), None, fused, name)


@pytest.mark.parametrize("w,h,maxw,maxh", [(640, 480, 160, 0),      # 4.0x: windows of 5 columns
                                           (641, 481, 128, 0),      # 5.0x: 6 columns (and 7 where the floor steps)
                                           (700, 500, 100, 0),      # 7.0x: 8 columns, the widest unrolled loop
                                           (900, 600, 75, 0),       # 12x: wider than the unrolled loops -> plain loop
                                           (640, 480, 0, 200),      # 2.4x from the height: 3-4 columns
                                           (333, 222, 160, 100)])   # 2.08x, ragged sizes
def test_scaled_window_widths(ip, orc, ctx, w, h, maxw, maxh):
    data = common.synth_cfa(w, h, seed=111)
    check(ip, orc, ctx, data, common.raw_params(cfa="GRBG"), {"maxwidth": maxw, "maxheight": maxh}, True, f"{w}x{h}")


@pytest.mark.parametrize("cfa", ["RGEB", common.XTRANS, "GMYE"])
def test_scaled_other_patterns(ip, orc, ctx, cfa):
    data = common.synth_cfa(600, 420, seed=121)
    matrix = common.CAM_TO_XYZ.copy()
    matrix[:, 3] = [0.05, 0.1, -0.02]
    check(ip, orc, ctx, data, common.raw_params(cfa=cfa, matrix=matrix, wb=[1.8, 1.0, 1.4, 1.1]), {"maxwidth": 150}, True, cfa)


@pytest.mark.parametrize("black,white", [(512.0, 16383.0), (0.0, 65535.0), (63.5, 4000.25), (1024.0, 1029.0)])
def test_level_mappings(ip, orc, ctx, black, white):
    """Levels for which the 3-instruction division is / is not exact for every sample (the host checks all 65536)."""
    data = common.synth_cfa(512, 96, seed=131)
    params = common.raw_params(black=black, white=white)
    check(ip, orc, ctx, data, params, None, True, f"levels {black}/{white}")
    check(ip, orc, ctx, data, params, {"maxwidth": 128}, True, f"levels {black}/{white} scaled")


def test_full_size_c3_xtrans_frame(ip, orc, ctx):
    """BASELINE config 3 at full size: 8256x5504 X-Trans -> 8-bit sRGB, whole frame against the oracle."""
    data = common.synth_cfa(8256, 5504, seed=common.SEED + 3)
    params = common.raw_params(cfa=common.XTRANS)
    want = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", params))
    got = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx).output_8bit().to_numpy()
    assert_bit_exact(got, want, "C3 frame")


def test_full_size_c5_frame_striped(ip, orc, ctx):
    """BASELINE config 5 at full size: 11648x8736 RGGB -> 8-bit sRGB.  The whole frame against the oracle, and the
    eight row stripes of an 8-GPU run (computed in turn on this GPU, each from its own rows + halo) against it."""
    from imagepipe_b200 import _capi
    from imagepipe_b200.sharded import plan_stripes, run_stripe_8bit
    w, h = 11648, 8736
    data = common.synth_cfa(w, h, seed=common.SEED + 5)
    params = common.raw_params()
    want = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", params))
    got = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx, on_device=True).output_8bit().to_numpy()
    assert_bit_exact(got, want, "C5 frame")
    del got
    dummy = ip.DeviceArray(64, ctx)
    p = ip.Pipeline.new_from_source(ip.ImageSource(_capi.SRC_RAW_U16, w, h, 1, dummy), ctx=ctx)
    common.fill_ipb_ops(p.ops, params)
    for lay in plan_stripes(p.ops, p.globals.settings, w, h, 8):
        assert lay.src_row1 - lay.src_row0 <= 1120 + 2  # 1092 rounded to the 32-row tile height, + halo
        rows = ip.DeviceArray.from_numpy(data[lay.src_row0:lay.src_row1], ctx)
        dst = ip.DeviceArray((lay.out_row1 - lay.out_row0) * w * 3, ctx)
        run_stripe_8bit(p, rows.ptr, lay, dst)
        assert_bit_exact(dst.to_numpy(np.uint8, (lay.out_row1 - lay.out_row0, w, 3)), want[lay.out_row0:lay.out_row1],
                         f"C5 stripe {lay.rank}")


@pytest.mark.parametrize("scale", [1.45, 2.5, 40.0])
def test_ratios_beyond_the_cube_root_table(ip, orc, ctx, scale):
    """XYZ ratios above 1.0 come from the cube-root table up to 1.5 and from the double-precision restatement of
    glibc's cbrtf beyond (the per-warp queue); a scaled camera matrix puts ratios on both sides of that border, and
    negative off-diagonals make some ratios negative.  Still one fused launch, still the oracle's bits."""
    data = common.synth_cfa(517, 203, seed=141)
    matrix = common.CAM_TO_XYZ * np.float32(scale)
    matrix[2, 1] = np.float32(-1.3 * scale)          # strongly negative: Z ratios below zero
    check(ip, orc, ctx, data, common.raw_params(matrix=matrix), None, True, f"matrix x{scale}")
