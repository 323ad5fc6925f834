"""GPU parity of the Lanczos extension (ipb_lanczos_resize) against its CPU statement oracle/lanczos.c: bit-exact
(same tap tables, same order of f32 operations).  Not a reference path: the reference has no Lanczos resampler."""
import ctypes as C

import numpy as np
import pytest

import common
from common import assert_bit_exact

pytestmark = pytest.mark.gpu


def orc_resize(orc, arr, nw, nh, a):
    bp = orc.buffer_from_numpy(arr)
    out = orc.lib().orc_lanczos_resize(bp, nw, nh, a)
    orc.lib().orc_buffer_free(bp)
    return orc.buffer_to_numpy(out)[0]


@pytest.mark.parametrize("shape,nw,nh,a", [((400, 600, 3), 150, 100, 3),     # 4x down
                                           ((333, 517, 3), 207, 133, 3),     # 2.5x, ragged
                                           ((64, 96, 1), 144, 96, 3),        # 1.5x up, one channel
                                           ((120, 160, 4), 160, 120, 2),     # same size, RGBE, a = 2
                                           ((1000, 1500, 3), 100, 67, 3),    # 15x down: wide taps
                                           ((37, 41, 3), 1, 1, 3),           # down to one pixel
                                           ((300, 4100, 3), 1025, 75, 4)])   # several column tiles, a = 4
def test_lanczos_matches_oracle(ip, orc, ctx, shape, nw, nh, a):
    rng = np.random.default_rng(5)
    arr = rng.uniform(-0.1, 1.2, shape).astype(np.float32)
    want = orc_resize(orc, arr, nw, nh, a)
    buf = ip.OpBuffer.from_numpy(arr, ctx=ctx)
    n0 = ctx.launch_count
    got = ip.lanczos_resize(buf, nw, nh, a)
    assert ctx.launch_count - n0 == 2
    assert (got.width, got.height, got.colors) == (nw, nh, shape[2])
    assert_bit_exact(got.to_numpy(), want, f"lanczos {shape} -> {nw}x{nh} a={a}")


def test_lanczos_after_the_pipeline(ip, orc, ctx):
    """The north star's order: the full pipe at full resolution, then the Lanczos reduction of its f32 output."""
    data = common.synth_cfa(1200, 800, seed=301)
    params = common.raw_params()
    full = orc.pipeline_run(orc.make_pipeline(data, "raw", params))
    want = orc_resize(orc, full, 300, 200, 3)
    p = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx)
    got = ip.lanczos_resize(p.run(), 300, 200)
    assert_bit_exact(got.to_numpy(), want, "pipeline + lanczos")


def test_lanczos_rejects_bad_arguments(ip, ctx):
    buf = ip.OpBuffer.from_numpy(np.zeros((8, 8, 3), np.float32), ctx=ctx)
    for nw, nh, a in ((0, 4, 3), (4, 4, 0), (4, 4, 9)):
        with pytest.raises(ip.IpbError):
            ip.lanczos_resize(buf, nw, nh, a)
