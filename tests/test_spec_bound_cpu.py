"""Host side of the speculative kernels without a GPU: ipb_spec_bound runs spec_build (ipb_spec_host.cu) — the certified
bound delta on |cheap linear value - reference linear value| per output channel — for a parameter set.  No device code."""
import ctypes as C

import numpy as np
import pytest

import common


def bound(ip, mufu=2.4e-7, **kw):
    ops = ip.PipelineOps()
    common.fill_ipb_ops(ops, common.raw_params(**kw))
    d = (C.c_float * 4)()
    rc = ip.lib().ipb_spec_bound(C.byref(ops), C.c_float(mufu), d)
    return rc, [float(v) for v in d]


@pytest.fixture(scope="module")
def ip():
    import imagepipe_b200
    return imagepipe_b200


def test_default_parameters(ip):
    rc, d = bound(ip)
    assert rc == 0
    assert d[0] == max(d[1:])                      # [0] is the largest channel
    assert all(5e-6 < v < 5e-5 for v in d[1:]), d  # a few 1e-5: thresholds are >= 3e-4 apart
    # red sees the largest matrix row of XYZ -> sRGB
    assert d[1] > d[2] and d[1] > d[3]


def test_bound_grows_with_the_measured_cube_root_error(ip):
    _, a = bound(ip, mufu=2.4e-7)
    _, b = bound(ip, mufu=1.0e-6)
    assert all(y > x for x, y in zip(a, b))


@pytest.mark.parametrize("kw", [dict(points=()), dict(points=((0.3, 0.2), (0.7, 0.9))), dict(wb=[1.0, 1.0, 1.0, 1.0]),
                                dict(black=64.0, white=4095.0), dict(exposure=-0.3)])
def test_usual_parameter_sets_are_certified(ip, kw):
    rc, d = bound(ip, **kw)
    assert rc == 0 and 0 < d[0] <= 8e-5, (rc, d)


def test_parameters_outside_the_measured_range_are_refused(ip):
    """XYZ ratios beyond what the XU-pipe cube root was measured on (matrix x 8): the speculative path is declined
    (IPB_ERR_UNSUPPORTED) and the pipeline takes the exact kernel."""
    rc, _ = bound(ip, matrix=common.CAM_TO_XYZ * np.float32(8.0))
    assert rc != 0


def test_scaled_division_check(ip):
    """ipb_scaled_division_check walks every tap of every window of a geometry: the reciprocal form of delta / skip
    (scaling.rs:94-98) equals IEEE division for BASELINE config 4 and for every geometry of a random sample (the form
    only fails for pathological divisors; the kernels divide the IEEE way when the check says 0); degenerate sizes: 0."""
    import random
    L = ip.lib()
    assert L.ipb_scaled_division_check(6000, 4000, 1500, 1000) == 1
    assert L.ipb_scaled_division_check(6000, 4000, 1, 1000) == 0
    rnd = random.Random(7)
    for _ in range(200):
        w = rnd.randint(64, 9000)
        nw = rnd.randint(max(2, w // 7 + 1), max(3, w // 2))
        h = rnd.randint(64, 6000)
        assert L.ipb_scaled_division_check(w, h, nw, max(2, h * nw // w)) == 1, (w, h, nw)
