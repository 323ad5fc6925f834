"""GPU parity of the fused raw->sRGB kernels (through Pipeline::run / output_8bit / output_16bit of the C ABI)
against the CPU oracle's op-by-op pipeline.  Bit-exact on f32, u8 and u16."""
import numpy as np
import pytest

import common
from common import assert_bit_exact

pytestmark = pytest.mark.gpu


def both(ip, orc, ctx, data, params, settings=None, on_device=False):
    po = orc.make_pipeline(data, "raw", params, settings)
    pg = common.make_ipb_pipeline(ip, data, "raw", params, settings, ctx=ctx, on_device=on_device)
    return po, pg


@pytest.mark.parametrize("cfa", ["RGGB", "BGGR", "GRBG", "GBRG"])
@pytest.mark.parametrize("shape", [(40, 64), (67, 301), (130, 523)])
def test_full_bayer_f32(ip, orc, ctx, cfa, shape):
    data = common.synth_cfa(shape[1], shape[0])
    po, pg = both(ip, orc, ctx, data, common.raw_params(cfa=cfa))
    want = orc.pipeline_run(po)
    got = pg.run().to_numpy()
    assert_bit_exact(got, want, f"fused full {cfa} {shape}")


@pytest.mark.parametrize("cfa", [common.XTRANS, "RGEB", "RGBE" * 4, "GMYE"])
def test_full_other_cfas(ip, orc, ctx, cfa):
    data = common.synth_cfa(277, 95, seed=5)
    matrix = common.CAM_TO_XYZ.copy()
    matrix[:, 3] = [0.05, 0.1, -0.02]  # give the E channel a weight
    po, pg = both(ip, orc, ctx, data, common.raw_params(cfa=cfa, matrix=matrix, wb=[1.8, 1.0, 1.4, 1.1]))
    assert_bit_exact(pg.run().to_numpy(), orc.pipeline_run(po), f"fused full {cfa}")


@pytest.mark.parametrize("crops", [(0, 0, 0, 0), (3, 5, 2, 7), (1, 0, 0, 1)])
def test_full_outputs_8_and_16(ip, orc, ctx, crops):
    data = common.smooth_cfa(410, 133)
    params = common.raw_params(cfa="GRBG", crops=crops)
    po, pg = both(ip, orc, ctx, data, params)
    want8 = orc.pipeline_output_8bit(po)
    got8 = pg.output_8bit()
    assert (got8.width, got8.height) == (want8.shape[1], want8.shape[0])
    assert_bit_exact(got8.to_numpy(), want8, "output_8bit")
    want16 = orc.pipeline_output_16bit(po)  # sets linear = true
    got16 = pg.output_16bit()
    assert_bit_exact(got16.to_numpy(), want16, "output_16bit")


def test_fused_equals_unfused_on_gpu(ip, ctx):
    data = common.synth_cfa(1031, 517, seed=11)
    pg = common.make_ipb_pipeline(ip, data, "raw", common.raw_params(), ctx=ctx)
    fused = pg.run().to_numpy()
    n0 = ctx.launch_count
    pg.set_fused(False)
    unfused = pg.run().to_numpy()
    assert ctx.launch_count - n0 >= 4  # gofloat, demosaic, to_lab + basecurve, from_lab + gamma (paired without a cache)
    assert_bit_exact(fused, unfused, "fused vs op-by-op")


@pytest.mark.parametrize("cfa,shape,maxw,maxh", [("RGGB", (200, 300), 75, 0), ("RGGB", (203, 311), 77, 0),
                                                 ("BGGR", (160, 240), 0, 20), (common.XTRANS, (180, 270), 60, 0),
                                                 ("RGGB", (120, 180), 100, 0)])
def test_scaled_pipeline(ip, orc, ctx, cfa, shape, maxw, maxh):
    """maxwidth/maxheight: scaled_demosaic (fused), or full()+scale_down_opbuf (op by op) below minscale."""
    data = common.synth_cfa(shape[1], shape[0], seed=21)
    st = {"maxwidth": maxw, "maxheight": maxh}
    po, pg = both(ip, orc, ctx, data, common.raw_params(cfa=cfa), st)
    want = orc.pipeline_run(po)
    got = pg.run().to_numpy()
    assert_bit_exact(got, want, f"scaled {cfa} {shape}")
    po, pg = both(ip, orc, ctx, data, common.raw_params(cfa=cfa), st)
    assert_bit_exact(pg.output_8bit().to_numpy(), orc.pipeline_output_8bit(po), "scaled output_8bit")


@pytest.mark.parametrize("rotation,fliph", [(1, False), (2, True), (3, False)])
def test_orientation_after_fused(ip, orc, ctx, rotation, fliph):
    data = common.synth_cfa(150, 90, seed=31)
    params = common.raw_params(rotation=rotation, fliph=fliph)
    po, pg = both(ip, orc, ctx, data, params, {"maxwidth": 100})
    want = orc.pipeline_output_8bit(po)
    got = pg.output_8bit()
    assert (got.width, got.height) == (want.shape[1], want.shape[0])
    assert_bit_exact(got.to_numpy(), want, "orientation")


def test_device_resident_source_and_destination(ip, orc, ctx):
    data = common.synth_cfa(512, 128, seed=41)
    po, pg = both(ip, orc, ctx, data, common.raw_params(), on_device=True)
    dst = ip.DeviceArray(512 * 128 * 3, ctx)
    img = pg.output_8bit(dst=dst)
    assert (img.width, img.height) == (512, 128)
    assert_bit_exact(dst.to_numpy(np.uint8, (128, 512, 3)), orc.pipeline_output_8bit(po), "device in/out")


@pytest.mark.parametrize("maxw", [0, 150])
def test_row_stripes_equal_whole_frame(ip, ctx, maxw):
    """Sharding property: stitching stripes (each given only the source rows stripe_rows() names) == whole frame."""
    h, w = 264, 600
    data = common.synth_cfa(w, h, seed=51)
    st = {"maxwidth": maxw}
    whole = common.make_ipb_pipeline(ip, data, "raw", common.raw_params(), st, ctx=ctx).output_8bit().to_numpy()
    oh = whole.shape[0]
    parts = []
    bounds = [0, oh // 3, oh // 3 + 1, (2 * oh) // 3, oh]
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        pg = common.make_ipb_pipeline(ip, data, "raw", common.raw_params(), st, ctx=ctx)
        s0, s1 = pg.stripe_rows(r0, r1)
        rows = np.ascontiguousarray(data[s0:s1])
        pg.set_stripe_source(ip.ImageSource.Raw(rows), s0, r0, r1)
        parts.append(pg.output_8bit_stripe(rows=r1 - r0, width=whole.shape[1]).to_numpy())
    assert_bit_exact(np.concatenate(parts, 0), whole, "stripes")


def test_stripe_missing_rows_is_an_error(ip, ctx):
    data = common.synth_cfa(64, 64)
    pg = common.make_ipb_pipeline(ip, data, "raw", common.raw_params(), ctx=ctx)
    pg.set_stripe_source(ip.ImageSource.Raw(np.ascontiguousarray(data[10:20])), 10, 10, 20)
    with pytest.raises(ip.IpbError):
        pg.output_8bit_stripe(rows=10, width=64)  # rows 9 and 20 (the halo) are missing


def test_synthetic_generator_matches_host(ip, ctx):
    d = ip.synth_cfa_u16(common.SEED + 3, 1000, 17, 9, ctx=ctx)
    assert_bit_exact(d.to_numpy(), common.synth_cfa(1000, 9, common.SEED + 3, row0=17), "synth")


def test_full_size_c2_frame(ip, orc, ctx):
    """BASELINE config 2 at full size: 6000x4000 RGGB -> 8-bit sRGB, whole frame against the oracle."""
    data = common.synth_cfa(6000, 4000)
    po, pg = both(ip, orc, ctx, data, common.raw_params())
    got = pg.output_8bit().to_numpy()
    want = orc.pipeline_output_8bit(po)
    assert_bit_exact(got, want, "C2 frame")


def test_full_size_c4_scaled(ip, orc, ctx):
    """BASELINE config 4 frame: 6000x4000 -> 1500x1000 through scaled_demosaic."""
    data = common.synth_cfa(6000, 4000, seed=common.SEED + 1)
    po, pg = both(ip, orc, ctx, data, common.raw_params(), {"maxwidth": 1500, "maxheight": 1000})
    got = pg.output_8bit()
    assert (got.width, got.height) == (1500, 1000)
    assert_bit_exact(got.to_numpy(), orc.pipeline_output_8bit(po), "C4 frame")


def test_xtrans_c3_rows(ip, orc, ctx):
    """BASELINE config 3 width (8256) on a 96-row band of the X-Trans frame."""
    data = common.synth_cfa(8256, 96, seed=common.SEED + 2)
    po, pg = both(ip, orc, ctx, data, common.raw_params(cfa=common.XTRANS))
    assert_bit_exact(pg.output_8bit().to_numpy(), orc.pipeline_output_8bit(po), "C3 band")


def test_gamma8_threshold_table_exhaustive(ip, ctx):
    """The fused 8-bit path's threshold table against the plain gamma lerp + output8bit on the device: every f32 in
    [0, 1] and a sample of all other bit patterns (negative, > 1, inf, NaN)."""
    import ctypes as C
    n = C.c_ulonglong(1)
    rc = ip.lib().ipb_selftest_gamma8(ctx.handle, C.byref(n))
    assert rc == 0, ip.lib().ipb_last_error(ctx.handle)
    assert n.value == 0


def test_gamma8_against_oracle(ip, orc, ctx):
    """gamma.rs:21 + color_conversions.rs:323-325 through the threshold table vs the oracle's OpGamma + output8bit,
    on random floats, the table knots, and values around every 8-bit step."""
    rng = np.random.default_rng(7)
    knots = (np.arange(8193, dtype=np.float32) / np.float32(8191)).astype(np.float32)
    vals = [rng.random(300000, dtype=np.float32), rng.normal(0.5, 0.6, 100000).astype(np.float32), knots,
            np.nextafter(knots, np.float32(0)), np.nextafter(knots, np.float32(2)),
            np.array([-0.0, 0.0, 1.0, 1.5, -1.0, np.inf, -np.inf, np.nan, 0.0031308, 0.00313], np.float32)]
    v = np.concatenate(vals).astype(np.float32)
    v = np.resize(v, (v.size // 3) * 3)
    L = orc.lib()
    st = orc.Settings()
    buf = orc.buffer_from_numpy(v.reshape(1, -1, 3))
    want_f = orc.buffer_to_numpy(L.orc_gamma_run(C_byref(st), buf))[0].reshape(-1)
    L.orc_buffer_free(buf)
    assert not np.isnan(want_f).any()  # OpGamma clamps NaN to 0 before the table (gamma.rs:21)
    want = np.clip(want_f * np.float32(256), 0, 255).astype(np.uint8)  # output8bit, color_conversions.rs:323-325
    got = np.empty(v.size, np.uint8)
    rc = ip.lib().ipb_gamma_pack_8bit(ctx.handle, v.ctypes.data, v.size, got.ctypes.data)
    assert rc == 0, ip.lib().ipb_last_error(ctx.handle)
    assert_bit_exact(got, want, "gamma8 vs oracle")


def C_byref(x):
    import ctypes as C
    return C.byref(x)


@pytest.mark.parametrize("cfa,shape,crops", [("RGGB", (70, 520), (0, 0, 0, 0)), ("GBRG", (53, 523), (0, 0, 0, 0)),
                                             ("BGGR", (96, 1032), (3, 5, 2, 7)), (common.XTRANS, (60, 528), (0, 0, 0, 0))])
def test_tma_and_plain_staging_agree(ip, orc, ctx, cfa, shape, crops):
    """The TMA-staged and the plain-load staging of the full-resolution kernel give the oracle's bytes; widths that
    are / are not a multiple of 8 samples exercise both the tensor-map path and its automatic fallback."""
    data = common.synth_cfa(shape[1], shape[0], seed=61)
    params = common.raw_params(cfa=cfa, crops=crops)
    want = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", params))
    for on_device in (False, True):
        for tma in (1, 0):
            pg = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx, on_device=on_device)
            assert ip.lib().ipb_pipeline_set_tma(pg.handle, tma) == 0
            assert_bit_exact(pg.output_8bit().to_numpy(), want, f"{cfa} tma={tma} dev={on_device}")


def test_natural_frame_mostly_in_table(ip, orc, ctx):
    """A smooth frame (few out-of-table XYZ ratios: the warp queue is mostly idle) and a clipped one (every pixel
    beyond white: the queue is full)."""
    for data in (common.smooth_cfa(777, 211), np.full((64, 512), 16383, np.uint16)):
        po, pg = both(ip, orc, ctx, data, common.raw_params())
        assert_bit_exact(pg.output_8bit().to_numpy(), orc.pipeline_output_8bit(po), "natural/clipped u8")
        po, pg = both(ip, orc, ctx, data, common.raw_params())
        assert_bit_exact(pg.run().to_numpy(), orc.pipeline_run(po), "natural/clipped f32")


@pytest.mark.parametrize("matrix_scale,wb", [(1000.0, common.WB), (1.0, [1e-9, 1.0, 1.5, 1.0]), (3e37, common.WB)])
def test_unbounded_parameters_run_op_by_op(ip, orc, ctx, matrix_scale, wb):
    """Outside the parameter bounds that make the fused kernels' constant divisions exact (ipb_host.cu
    fused_params_bounded) Pipeline::run must fall back to one kernel per op (IEEE division) and still match."""
    data = common.synth_cfa(301, 67, seed=71)
    params = common.raw_params(matrix=common.CAM_TO_XYZ * np.float32(matrix_scale), wb=wb)
    po, pg = both(ip, orc, ctx, data, params)
    n0 = ctx.launch_count
    got = pg.run().to_numpy()
    assert ctx.launch_count - n0 >= 4   # gofloat, demosaic, to_lab + basecurve, from_lab + gamma
    assert_bit_exact(got, orc.pipeline_run(po), "op-by-op fallback")
