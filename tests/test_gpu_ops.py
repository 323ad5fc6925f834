"""GPU parity, op by op: every ImageOp::run of the C ABI against the CPU oracle on the same seeded inputs.
Bar: bit-exact f32 (the device code is compiled -fmad=false and shares the host-built tables)."""
import numpy as np
import pytest

import common
from common import assert_bit_exact

pytestmark = pytest.mark.gpu


def orc_buf(orc, arr, mono=False):
    return orc.buffer_from_numpy(arr, mono)


def run_orc(orc, fn, *args):
    bp = fn(*args)
    assert bp, "oracle returned NULL"
    return orc.buffer_to_numpy(bp)[0]


@pytest.fixture(scope="module")
def rgbe(orc):
    rng = np.random.default_rng(7)
    a = rng.uniform(-0.05, 1.3, (61, 77, 4)).astype(np.float32)
    a[0, 0] = [0, 0, 0, 0]
    a[0, 1] = [1, 1, 1, 1]
    a[0, 2] = [-0.0, 0.5, 2.0, 0.0]
    a[0, 3] = [np.nan, 0.2, 0.3, 0.0]
    a[..., 3] = 0.0
    return a


@pytest.mark.parametrize("kind,cpp,is_cfa", [("u16", 1, True), ("u16", 1, False), ("u16", 3, True),
                                             ("f32", 1, True), ("f32", 1, False), ("f32", 3, False)])
def test_gofloat_raw(ip, orc, ctx, kind, cpp, is_cfa):
    import ctypes as C
    rng = np.random.default_rng(3)
    shape = (37, 53) if cpp == 1 else (37, 53, cpp)
    data = rng.integers(0, 16384, shape).astype(np.uint16) if kind == "u16" else \
        rng.uniform(0, 1.2, shape).astype(np.float32)
    params = common.raw_params(crops=(2, 3, 1, 4), black=0.03 if kind == "f32" else 512.0,
                               white=0.97 if kind == "f32" else 16383.0)
    params["gofloat"]["is_cfa"] = is_cfa
    params["gofloat"]["blacklevels"] = [params["gofloat"]["blacklevels"][0] * s for s in (1, 1.1, 0.9, 1)]
    src, keep = orc.make_source(data, "raw", cpp)
    ops = orc.Ops()
    orc.fill_ops(ops, params)
    want, mono = orc.buffer_to_numpy(orc.lib().orc_gofloat_run(C.byref(ops.gofloat), C.byref(src)))
    p = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx)
    got = p.ops.gofloat.run(p.globals)
    assert got.monochrome == mono
    assert_bit_exact(got.to_numpy(), want, "gofloat")


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_gofloat_other(ip, orc, ctx, dtype):
    import ctypes as C
    rng = np.random.default_rng(4)
    data = rng.integers(0, np.iinfo(dtype).max + 1, (33, 41, 3)).astype(dtype)
    src, keep = orc.make_source(data, "rgb")
    ops = orc.Ops()
    want, _ = orc.buffer_to_numpy(orc.lib().orc_gofloat_run(C.byref(ops.gofloat), C.byref(src)))
    p = common.make_ipb_pipeline(ip, data, "rgb", ctx=ctx)
    assert_bit_exact(p.ops.gofloat.run(p.globals).to_numpy(), want, "gofloat other")


@pytest.mark.parametrize("cfa", ["RGGB", "BGGR", "GRBG", "GBRG", "RGEB", common.XTRANS, "RGBE" * 4])
@pytest.mark.parametrize("shape", [(10, 10), (47, 61)])
def test_demosaic_full(ip, orc, ctx, cfa, shape):
    import ctypes as C
    rng = np.random.default_rng(5)
    a = rng.uniform(-0.04, 1.0, shape).astype(np.float32)
    c = orc.Cfa()
    assert orc.lib().orc_cfa_new(C.byref(c), cfa.encode()) == 0
    ob = orc_buf(orc, a)
    want = run_orc(orc, orc.lib().orc_demosaic_full, C.byref(c), ob)
    orc.lib().orc_buffer_free(ob)
    op = ip.OpDemosaic()
    op.cfa = cfa.encode()
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    g.settings.demosaic_width, g.settings.demosaic_height = shape[1], shape[0]
    got = op.run(g, ip.OpBuffer.from_numpy(a, ctx=ctx))
    assert got.colors == 4
    assert_bit_exact(got.to_numpy(), want, f"demosaic full {cfa}")


@pytest.mark.parametrize("cfa,src,dst", [("RGGB", (64, 96), (16, 24)), ("RGGB", (100, 150), (33, 50)),
                                         (common.XTRANS, (96, 120), (24, 30)), ("GBRG", (57, 83), (11, 16)),
                                         ("RGGB", (60, 90), (40, 60))])
def test_demosaic_scaled_branches(ip, orc, ctx, cfa, src, dst):
    """OpDemosaic::run branch selection (demosaic.rs:41-60): scaled_demosaic and full()+scale_down_opbuf."""
    import ctypes as C
    rng = np.random.default_rng(6)
    a = rng.uniform(0, 1.0, src).astype(np.float32)
    dop = orc.Demosaic()
    dop.cfa = cfa.encode()
    st = orc.Settings()
    st.demosaic_width, st.demosaic_height = dst[1], dst[0]
    ob = orc_buf(orc, a)
    want = run_orc(orc, orc.lib().orc_demosaic_run, C.byref(dop), C.byref(st), ob)
    orc.lib().orc_buffer_free(ob)
    op = ip.OpDemosaic()
    op.cfa = cfa.encode()
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    g.settings.demosaic_width, g.settings.demosaic_height = dst[1], dst[0]
    got = op.run(g, ip.OpBuffer.from_numpy(a, ctx=ctx))
    assert_bit_exact(got.to_numpy(), want, f"demosaic scaled {cfa} {src}->{dst}")


def test_demosaic_4ch_passthrough_and_scale(ip, orc, ctx, rgbe):
    import ctypes as C
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    op = ip.OpDemosaic()
    buf = ip.OpBuffer.from_numpy(rgbe, ctx=ctx)
    g.settings.demosaic_width, g.settings.demosaic_height = rgbe.shape[1], rgbe.shape[0]
    assert op.run(g, buf).same_arc(buf)  # demosaic.rs:41-43 returns the same Arc
    g.settings.demosaic_width, g.settings.demosaic_height = 25, 20
    ob = orc_buf(orc, rgbe)
    want = run_orc(orc, orc.lib().orc_scale_down_opbuf, ob, 25, 20)
    orc.lib().orc_buffer_free(ob)
    assert_bit_exact(op.run(g, buf).to_numpy(), want, "scale_down_opbuf")


@pytest.mark.parametrize("mono", [False, True])
def test_tolab(ip, orc, ctx, rgbe, mono):
    import ctypes as C
    params = common.raw_params()
    ops = orc.Ops()
    orc.fill_ops(ops, params)
    ob = orc_buf(orc, rgbe, mono)
    want = run_orc(orc, orc.lib().orc_tolab_run, C.byref(ops.tolab), ob)
    orc.lib().orc_buffer_free(ob)
    p_ops = ip.PipelineOps()
    common.fill_ipb_ops(p_ops, params)
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    got = p_ops.tolab.run(g, ip.OpBuffer.from_numpy(rgbe, monochrome=mono, ctx=ctx))
    assert got.colors == 3 and got.monochrome == mono
    assert_bit_exact(got.to_numpy(), want, "to_lab")


@pytest.mark.parametrize("scale", [1.3, 7.0, 1000.0, 3e37])
def test_lab_transfer_above_one(ip, orc, ctx, scale):
    """XYZ ratios far above 1.0 take the reference's analytic branch: v.cbrt() == glibc cbrtf, which the device
    restates (double-precision Halley step).  The matrix is scaled so the ratios cover (1, scale] (and +inf)."""
    import ctypes as C
    rng = np.random.default_rng(5)
    rgbe = rng.random((37, 211, 4), dtype=np.float32)
    rgbe[0, :8, :3] = 1.0
    params = common.raw_params(matrix=common.CAM_TO_XYZ * np.float32(scale), wb=[1.0, 1.0, 1.0, 1.0])
    ops = orc.Ops()
    orc.fill_ops(ops, params)
    ob = orc_buf(orc, rgbe, False)
    want = run_orc(orc, orc.lib().orc_tolab_run, C.byref(ops.tolab), ob)
    orc.lib().orc_buffer_free(ob)
    p_ops = ip.PipelineOps()
    common.fill_ipb_ops(p_ops, params)
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    got = p_ops.tolab.run(g, ip.OpBuffer.from_numpy(rgbe, ctx=ctx))
    assert_bit_exact(got.to_numpy(), want, f"to_lab x{scale}")


@pytest.fixture(scope="module")
def lab(orc):
    rng = np.random.default_rng(8)
    a = rng.uniform(-0.1, 1.1, (45, 67, 3)).astype(np.float32)
    a[0, 0] = [0, 0, 0]
    a[0, 1] = [1, 1, 1]
    a[0, 2] = [0.5, 0.6, np.nan]
    return a


@pytest.mark.parametrize("points,exposure", [([(0.5, 0.6)], 0.0), ([], 0.5), ([(0.2, 0.1), (0.7, 0.9)], -0.3),
                                             ([(0.0, 0.2)], 0.0), ([(1.0, 0.8)], 0.0),
                                             ([(0.1, 0.3), (0.3, 0.2), (0.6, 0.7), (0.9, 0.85)], 0.0)])
def test_basecurve(ip, orc, ctx, lab, points, exposure):
    import ctypes as C
    params = common.raw_params(points=points, exposure=exposure)
    ops = orc.Ops()
    orc.fill_ops(ops, params)
    ob = orc_buf(orc, lab)
    want = run_orc(orc, orc.lib().orc_basecurve_run, C.byref(ops.basecurve), ob)
    orc.lib().orc_buffer_free(ob)
    p_ops = ip.PipelineOps()
    common.fill_ipb_ops(p_ops, params)
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    got = p_ops.basecurve.run(g, ip.OpBuffer.from_numpy(lab, ctx=ctx))
    assert_bit_exact(got.to_numpy(), want, f"basecurve {points}")


def test_basecurve_noop_returns_same_arc(ip, ctx, lab):
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    buf = ip.OpBuffer.from_numpy(lab, ctx=ctx)
    op = ip.OpBaseCurve()
    assert op.run(g, buf).same_arc(buf)  # curves.rs:34-36


def test_fromlab_gamma_pack(ip, orc, ctx, lab):
    ob = orc_buf(orc, lab)
    rgb = run_orc(orc, orc.lib().orc_fromlab_run, ob)
    orc.lib().orc_buffer_free(ob)
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    got = ip.OpFromLab().run(g, ip.OpBuffer.from_numpy(lab, ctx=ctx))
    assert_bit_exact(got.to_numpy(), rgb, "from_lab")
    import ctypes as C
    st = orc.Settings()
    ob = orc_buf(orc, rgb)
    gam = run_orc(orc, orc.lib().orc_gamma_run, C.byref(st), ob)
    orc.lib().orc_buffer_free(ob)
    gbuf = ip.OpGamma().run(g, got)
    assert_bit_exact(gbuf.to_numpy(), gam, "gamma")
    g.settings.linear = 1
    assert ip.OpGamma().run(g, got).same_arc(got)  # gamma.rs:17-18
    # pack loops (pipeline.rs:408-414, 455-461)
    flat = gam.reshape(-1)
    want8 = np.array([orc.lib().orc_output8bit(float(v)) for v in flat[:3000]], np.uint8)
    want16 = np.array([orc.lib().orc_output16bit(float(v)) for v in flat[:3000]], np.uint16)
    out8 = np.empty(flat.size, np.uint8)
    out16 = np.empty(flat.size, np.uint16)
    from imagepipe_b200 import _capi
    _capi.check(ctx.handle, ip.lib().ipb_pack_8bit(ctx.handle, gbuf.handle, out8.ctypes.data, 0))
    _capi.check(ctx.handle, ip.lib().ipb_pack_16bit(ctx.handle, gbuf.handle, out16.ctypes.data, 0))
    assert_bit_exact(out8[:3000], want8, "pack8")
    assert_bit_exact(out16[:3000], want16, "pack16")


@pytest.mark.parametrize("rotation", [0, 1, 2, 3])
@pytest.mark.parametrize("fliph", [False, True])
@pytest.mark.parametrize("flipv", [False, True])
def test_transform_bit_exact(ip, orc, ctx, rotation, fliph, flipv):
    import ctypes as C
    rng = np.random.default_rng(9)
    a = rng.uniform(0, 1, (37, 70, 3)).astype(np.float32)
    top = orc.Transform(rotation, int(fliph), int(flipv))
    ob = orc_buf(orc, a)
    bp = orc.lib().orc_transform_run(C.byref(top), ob)
    same = C.addressof(bp.contents) == C.addressof(ob.contents)
    want = orc.buffer_to_numpy(bp, free=not same)[0]
    orc.lib().orc_buffer_free(ob)
    buf = ip.OpBuffer.from_numpy(a, ctx=ctx)
    got = ip.OpTransform(rotation, int(fliph), int(flipv)).run(ip.PipelineGlobals.mock(16, 16, ctx=ctx), buf)
    assert got.same_arc(buf) == same
    assert_bit_exact(got.to_numpy(), want, "transform")


@pytest.mark.parametrize("crop,rot", [((0.1, 0, 0, 0), 0.0), ((0.1, 0.05, 0.2, 0.15), 0.0), ((0, 0, 0, 0), 0.5),
                                      ((0.05, 0.1, 0.0, 0.1), 0.3), ((0, 0, 0, 0), 1.0)])
@pytest.mark.parametrize("colors", [3, 4])
def test_rotatecrop(ip, orc, ctx, crop, rot, colors):
    import ctypes as C
    rng = np.random.default_rng(10)
    a = rng.uniform(0, 1, (90, 120, colors)).astype(np.float32)
    rop = orc.RotateCrop(crop[0], crop[1], crop[2], crop[3], rot, 1.0, 0, 0, 0)
    ob = orc_buf(orc, a)
    want = run_orc(orc, orc.lib().orc_rotatecrop_run, C.byref(rop), ob)
    orc.lib().orc_buffer_free(ob)
    op = ip.OpRotateCrop.empty()
    op.crop_top, op.crop_right, op.crop_bottom, op.crop_left, op.rotation = (*crop, rot)
    got = op.run(ip.PipelineGlobals.mock(16, 16, ctx=ctx), ip.OpBuffer.from_numpy(a, ctx=ctx))
    assert_bit_exact(got.to_numpy(), want, "rotatecrop")


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_scale_down_srgb(ip, orc, ctx, dtype):
    rng = np.random.default_rng(11)
    a = rng.integers(0, np.iinfo(dtype).max + 1, (75, 101, 3)).astype(dtype)
    want = np.empty((20, 27, 3), dtype)
    fn = orc.lib().orc_scale_down_srgb if dtype == np.uint8 else orc.lib().orc_scale_down_srgb16
    fn(a.ctypes.data, 101, 75, 27, 20, want.ctypes.data)
    assert_bit_exact(ip.scale_down_srgb(a, 27, 20, ctx=ctx), want, "scale_down_srgb")


def test_wrong_colors_is_an_error_not_an_abort(ip, ctx):
    g = ip.PipelineGlobals.mock(16, 16, ctx=ctx)
    buf = ip.OpBuffer.new(8, 8, 4, ctx=ctx)
    with pytest.raises(ip.IpbError) as e:
        ip.OpTransform(1, 0, 0).run(g, buf)  # transform.rs:88 assert_eq!(buf.colors, 3)
    assert e.value.code == 2
    with pytest.raises(ip.IpbError):
        ip.OpFromLab().run(g, buf)
    op = ip.OpDemosaic()
    op.cfa = b"RGGBX"
    with pytest.raises(ip.IpbError) as e:
        op.run(g, ip.OpBuffer.new(8, 8, 1, ctx=ctx))
    assert e.value.code == 3
