"""The reference's own known-answer tests, run against the CUDA path through the C ABI (the same asserts as
tests/test_oracle_kats.py makes of the oracle; every case cites the reference test it restates)."""
import numpy as np
import pytest

import common
from common import assert_bit_exact
from test_oracle_kats import (F_ROWS, MAXSIZE_CASES, OP_FIELDS, ORIENT_GOLD, all_colors_8bit, block_16bit, rgb_str)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fast", [True, False])
def test_roundtrip_8bit_all_colors(ip, ctx, fast):
    """tests/roundtrip_test.rs:4-35 — every (R,G,B) u8 through the whole pipeline (fast path, and op by op through
    gofloat(run_other) -> to_lab -> basecurve -> from_lab -> gamma -> pack) comes back unchanged."""
    img = all_colors_8bit()
    p = common.make_ipb_pipeline(ip, img, "rgb", settings={"use_fastpath": fast}, ctx=ctx)
    n0 = ctx.launch_count
    out = p.output_8bit().to_numpy()
    assert_bit_exact(out, img, f"8-bit round trip, fastpath={fast}")
    if not fast:
        assert ctx.launch_count - n0 >= 4  # really went through the ops (gofloat, to_lab + basecurve, from_lab + gamma, pack)


@pytest.mark.parametrize("fast", [True, False])
def test_roundtrip_16bit_every_block(ip, ctx, fast):
    """tests/roundtrip_test.rs:37-84 — all blocks of the strided 16-bit colour grid through output_16bit."""
    _, nblocks = block_16bit(0)
    for blk in range(nblocks):
        img, _ = block_16bit(blk)
        p = common.make_ipb_pipeline(ip, img, "rgb", settings={"use_fastpath": fast}, ctx=ctx)
        assert_bit_exact(p.output_16bit().to_numpy(), img, f"16-bit round trip block {blk}, fastpath={fast}")


@pytest.mark.parametrize("settings,params,size", MAXSIZE_CASES)
def test_maxsize_kats(ip, ctx, settings, params, size):
    """tests/maxsize_test.rs:32-90"""
    img = np.zeros((64, 128, 3), np.uint8)
    for fast in (True, False):
        p = common.make_ipb_pipeline(ip, img, "rgb", params, dict(settings, use_fastpath=fast), ctx=ctx)
        out8 = p.output_8bit()
        assert (out8.width, out8.height) == size
        p = common.make_ipb_pipeline(ip, img, "rgb", params, dict(settings, use_fastpath=fast), ctx=ctx)
        out16 = p.output_16bit()
        assert (out16.width, out16.height) == size


def test_curves_kats(ip, ctx):
    """curves.rs:164-189 through the device spline kernel"""
    f = ip.SplineFunc([], ctx=ctx)
    assert f.interpolate(0.0) == 0.0 and f.interpolate(1.0) == 1.0          # extremes
    assert f.interpolate(1.5) == 1.0 and f.interpolate(-0.2) == 0.0         # saturates
    assert ip.SplineFunc([(0.0, 0.2)], ctx=ctx).interpolate(0.0) == float(np.float32(0.2))   # high_blackpoint
    assert ip.SplineFunc([(1.0, 0.8)], ctx=ctx).interpolate(1.0) == float(np.float32(0.8))   # low_whitepoint


@pytest.mark.parametrize("name", sorted(ORIENT_GOLD))
def test_transform_orientation_kats(ip, ctx, name):
    """transform.rs:167-278: the eight golden bitmaps, via rotate_buffer and via OpTransform's fields"""
    src = ip.OpBuffer.from_rgb_str_vec(F_ROWS, ctx=ctx)
    want = rgb_str(ORIENT_GOLD[name])
    got = ip.rotate_buffer(src, name)
    assert_bit_exact(got.to_numpy(), want, f"rotate_buffer {name}")
    if name != "Transverse":  # OpTransform::new maps Transverse onto fields whose run() gives Transpose (reference quirk)
        op = ip.OpTransform(*OP_FIELDS[name])
        out = op.run(ip.PipelineGlobals.mock(16, 16, ctx=ctx), src)
        assert_bit_exact(out.to_numpy(), want, f"OpTransform {name}")
        if name in ("Normal", "Unknown"):
            assert out.same_arc(src)  # transform.rs:68-69 returns the same Arc


def test_scaling_noop(ip, ctx):
    """scaling.rs:188-203"""
    w = h = 150
    data = np.arange(w * h * 3, dtype=np.uint32).astype(np.uint16).reshape(h, w, 3)
    assert_bit_exact(ip.scale_down_srgb(data, w, h, ctx=ctx), data, "scale_down_srgb16 identity")


@pytest.mark.parametrize("crops,size,first", [
    (dict(crop_top=0.1), (100, 90), 100 * 10 * 3), (dict(crop_bottom=0.1), (100, 90), 0),
    (dict(crop_top=0.1, crop_bottom=0.1), (100, 80), 100 * 10 * 3), (dict(crop_left=0.1), (90, 100), 10 * 3),
    (dict(crop_right=0.1), (90, 100), 0), (dict(crop_left=0.1, crop_right=0.1), (80, 100), 10 * 3),
    (dict(crop_left=0.1, crop_right=0.1, crop_top=0.1, crop_bottom=0.1), (80, 80), 100 * 10 * 3 + 10 * 3),
    (dict(rotation=0.5), (141, 141), None), (dict(rotation=1.0), (100, 100), None)])
def test_rotatecrop_kats(ip, ctx, crops, size, first):
    """rotatecrop.rs:185-271"""
    a = np.arange(100 * 100 * 3, dtype=np.float32).reshape(100, 100, 3)
    buf = ip.OpBuffer.from_numpy(a, ctx=ctx)
    op = ip.OpRotateCrop.empty()
    for k, v in crops.items():
        setattr(op, k, v)
    got = op.run(ip.PipelineGlobals.mock(100, 100, ctx=ctx), buf)
    assert (got.width, got.height) == size
    if first is not None:
        assert got.to_numpy().reshape(-1)[0] == a.reshape(-1)[first]
