"""Pipeline::run(Some(&cache)) on the GPU (pipeline.rs:340-372): the device-resident LRU of op outputs, re-entry at
the first op whose parameters changed, and equality with the uncached / fused result and the oracle."""
import numpy as np
import pytest

import common
from common import assert_bit_exact

pytestmark = pytest.mark.gpu
OPS = ["gofloat", "demosaic", "rotatecrop", "to_lab", "basecurve", "from_lab", "gamma", "transform"]


def test_cached_run_reenters_at_the_first_changed_op(ip, orc, ctx):
    data = common.synth_cfa(520, 260, seed=201)
    params = common.raw_params()
    p = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx, on_device=True)
    cache = ip.Pipeline.new_cache(1 << 30, ctx)
    want = orc.pipeline_run(orc.make_pipeline(data, "raw", params))
    first = p.run(cache)
    assert p.last_run_info() == (0, 8)                     # nothing cached: every op ran
    assert cache.entries == 8
    assert_bit_exact(first.to_numpy(), want, "cached run, cold")
    again = p.run(cache)
    assert p.last_run_info() == (8, 0)                     # the final buffer itself came from the cache
    assert again.same_arc(first)
    # a new base curve: gofloat .. to_lab come from the cache, basecurve .. transform run
    p.ops.basecurve.set_points([(0.3, 0.2), (0.7, 0.9)])
    params2 = common.raw_params(points=((0.3, 0.2), (0.7, 0.9)))
    got = p.run(cache)
    assert p.last_run_info() == (OPS.index("basecurve"), 4)
    assert_bit_exact(got.to_numpy(), orc.pipeline_run(orc.make_pipeline(data, "raw", params2)), "after curve change")
    # back to the first curve: its final buffer is still cached
    p.ops.basecurve.set_points([(0.5, 0.6)])
    assert p.run(cache).same_arc(first) and p.last_run_info() == (8, 0)
    # white balance: re-entry at to_lab
    p.ops.tolab.wb_coeffs[0] = 1.7
    p.run(cache)
    assert p.last_run_info() == (OPS.index("to_lab"), 5)
    # a setting every hash depends on (pipeline.rs:346): everything runs again
    p.globals.settings.maxwidth = 130
    small = p.run(cache)
    assert p.last_run_info() == (0, 8) and (small.width, small.height) == (130, 65)


def test_cached_outputs_match_uncached(ip, orc, ctx):
    data = common.synth_cfa(300, 200, seed=211)
    params = common.raw_params(cfa="GBRG", rotation=1)
    cache = ip.Pipeline.new_cache(1 << 28, ctx)
    p = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx)
    want8 = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", params))
    assert_bit_exact(p.output_8bit(cache).to_numpy(), want8, "output_8bit(cache)")
    assert_bit_exact(p.output_8bit().to_numpy(), want8, "output_8bit()")
    want16 = orc.pipeline_output_16bit(orc.make_pipeline(data, "raw", params))
    assert_bit_exact(p.output_16bit(cache).to_numpy(), want16, "output_16bit(cache)")   # linear: new settings hash
    assert p.last_run_info() == (0, 8)


def test_cache_is_size_bounded_lru(ip, ctx):
    data = common.synth_cfa(256, 128, seed=221)
    px = 256 * 128 * 4
    # room for the 1-channel gofloat output and two 3-channel buffers, not for the 4-channel demosaic output as well
    cache = ip.Pipeline.new_cache(px * 7, ctx)
    p = common.make_ipb_pipeline(ip, data, "raw", common.raw_params(), ctx=ctx)
    ref = p.run().to_numpy()
    got = p.run(cache)
    assert cache.bytes <= px * 7 and 0 < cache.entries < 8
    assert_bit_exact(got.to_numpy(), ref, "bounded cache")
    assert_bit_exact(p.run(cache).to_numpy(), ref, "bounded cache, second run")
    cache.clear()
    assert cache.entries == 0 and cache.bytes == 0
    tiny = ip.Pipeline.new_cache(16, ctx)   # nothing fits: still correct, nothing stored
    assert_bit_exact(p.run(tiny).to_numpy(), ref, "cache too small for anything")
    assert tiny.entries == 0


def test_two_pipelines_share_a_cache(ip, ctx):
    a = common.synth_cfa(200, 120, seed=231)
    b = common.synth_cfa(200, 120, seed=232)
    cache = ip.Pipeline.new_cache(1 << 28, ctx)
    pa = common.make_ipb_pipeline(ip, a, "raw", common.raw_params(), ctx=ctx, on_device=True)
    pb = common.make_ipb_pipeline(ip, b, "raw", common.raw_params(), ctx=ctx, on_device=True)
    ra, rb = pa.run(cache).to_numpy(), pb.run(cache).to_numpy()
    assert pb.last_run_info() == (0, 8)                     # same parameters, other pixels: no false hit
    assert (ra != rb).any()
    assert_bit_exact(pa.run(cache).to_numpy(), ra, "pipeline a again")
    assert pa.last_run_info() == (8, 0)


@pytest.mark.parametrize("depth", [np.uint8, np.uint16])
@pytest.mark.parametrize("maxwidth", [0, 97])
def test_other_source_with_a_cache_takes_the_fast_path_first(ip, orc, ctx, depth, maxwidth):
    """pipeline.rs:381 / :428: for a non-raw source with default ops the fast path comes before run(cache), with or
    without a cache — output_16bit(cache) is the gamma-encoded raster (not a linear run), output_8bit(cache) with a
    size limit goes through scale_down_srgb on the integers.  Cached and uncached calls give the same image."""
    rng = np.random.default_rng(7)
    img = rng.integers(0, np.iinfo(depth).max + 1, (120, 200, 3)).astype(depth)
    st = {"maxwidth": maxwidth} if maxwidth else None
    cache = ip.Pipeline.new_cache(1 << 28, ctx)
    for bits, call, ocall in ((8, "output_8bit", orc.pipeline_output_8bit), (16, "output_16bit", orc.pipeline_output_16bit)):
        want = ocall(orc.make_pipeline(img, "rgb", None, st))
        p = common.make_ipb_pipeline(ip, img, "rgb", None, st, ctx=ctx)
        assert_bit_exact(getattr(p, call)().to_numpy(), want, f"{call}() fast path")
        assert_bit_exact(getattr(p, call)(cache).to_numpy(), want, f"{call}(cache) fast path")
        # without the fast path both go through the pipeline, cached or not, and still agree with the oracle's slow path
        st_slow = dict(st or {}, use_fastpath=0)
        want_slow = ocall(orc.make_pipeline(img, "rgb", None, st_slow))
        ps = common.make_ipb_pipeline(ip, img, "rgb", None, st_slow, ctx=ctx)
        assert_bit_exact(getattr(ps, call)(cache).to_numpy(), want_slow, f"{call}(cache) slow path")


def test_set_source_invalidates_the_cache_for_a_refilled_buffer(ip, orc, ctx):
    """A ring buffer refilled with the next frame keeps its address: set_source must still start a new hash chain."""
    params = common.raw_params()
    a, b = common.synth_cfa(256, 128, seed=301), common.synth_cfa(256, 128, seed=302)
    d = ip.DeviceArray.from_numpy(a, ctx)
    src = ip.ImageSource.Raw(d, 256, 128)
    p = ip.Pipeline.new_from_source(src, ctx=ctx)
    common.fill_ipb_ops(p.ops, params)
    cache = ip.Pipeline.new_cache(1 << 28, ctx)
    assert_bit_exact(p.run(cache).to_numpy(), orc.pipeline_run(orc.make_pipeline(a, "raw", params)), "frame a")
    ip.lib().ipb_device_upload(ctx.handle, d.ptr, b.ctypes.data, b.nbytes)   # same device buffer, next frame
    p.set_source(src)
    assert_bit_exact(p.run(cache).to_numpy(), orc.pipeline_run(orc.make_pipeline(b, "raw", params)), "frame b")
    assert p.last_run_info() == (0, 8)


def test_stripe_rows_query_leaves_settings_alone(ip, ctx):
    data = common.synth_cfa(256, 128, seed=311)
    p = common.make_ipb_pipeline(ip, data, "raw", common.raw_params(), ctx=ctx, on_device=True)
    p.output_16bit()
    assert p.globals.settings.linear == 1          # sticky state of the last output call (pipeline.rs:452)
    p.stripe_rows(0, 64)
    assert p.globals.settings.linear == 1


CURVE_CASES = [((0.5, 0.6),), ((0.3, 0.2), (0.7, 0.9)), ((0.6, 0.5), (0.3, 0.4)), (), ((0.2, 0.6), (0.5, 0.3), (0.8, 0.7))]


@pytest.mark.parametrize("points", CURVE_CASES, ids=["default", "two knots", "unsorted", "passthrough", "non-monotone"])
@pytest.mark.parametrize("matrix_scale", [1.0, 3e37])
def test_paired_ops_equal_the_separate_ops(ip, ctx, points, matrix_scale):
    """Pipeline::run without a cache runs to_lab + basecurve and from_lab + gamma as one kernel each (k_tolab<1|2>,
    k_fromlab_gamma); with a cache every op runs on its own (k_tolab<0>, k_basecurve, k_fromlab, k_gamma) because each
    result is kept.  Same bits either way: sorted and unsorted knots, a pass-through curve, finite and overflowing
    (inf / NaN) pixel values."""
    data = common.synth_cfa(333, 91, seed=77)
    params = common.raw_params(points=points, matrix=common.CAM_TO_XYZ * np.float32(matrix_scale))
    p = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx)
    p.set_fused(False)
    n0 = ctx.launch_count
    paired = p.run().to_numpy()
    n_paired = ctx.launch_count - n0
    cache = ip.Pipeline.new_cache(1 << 28, ctx)
    n0 = ctx.launch_count
    separate = p.run(cache).to_numpy()
    n_separate = ctx.launch_count - n0
    assert n_paired < n_separate, (n_paired, n_separate)
    assert_bit_exact(paired, separate, f"paired vs separate, curve {points}, matrix x {matrix_scale}")
