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


# ---------------------------------------------------------------- the 8-bit stage of the cheap pass, emulated on the CPU

def _f32_fma(a, b, c, down=False):
    """f32 fma of f32 arrays through float64 (the product of two 24-bit significands is exact in 53 bits; the sums met
    here are exact too), rounded to nearest or towards -inf."""
    t = a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)
    f = t.astype(np.float32)
    if down:
        f = np.where(f.astype(np.float64) > t, np.nextafter(f, np.float32(-np.inf)), f)
    return f


def _spec_tables(ip, delta_override=0.0, **kw):
    ops = ip.PipelineOps()
    common.fill_ipb_ops(ops, common.raw_params(**kw))
    g8a = np.zeros(8192, np.uint32)
    thr = np.zeros(255, np.float32)
    one = (C.c_float * 3)()
    wmul = (C.c_uint32 * 3)()
    amb = C.c_uint32()
    d = (C.c_float * 4)()
    rc = ip.lib().ipb_spec_tables(C.byref(ops), C.c_float(2.4e-7), C.c_float(delta_override), g8a.ctypes.data, thr.ctypes.data,
                                  one, wmul, C.byref(amb), d)
    assert rc == 0
    return g8a, thr, [float(v) for v in one], [int(v) for v in wmul], int(amb.value), [float(v) for v in d]


def _cheap_bytes(v, g8a, one_c, wmul_c, amb_t):
    """chain_pair's last stage (ipb_spec.cu): u = fma(v, 1 - 2^-13, one_c), k = fma_rd(v, 32764 / 2^23, 1), table word
    addressed by bits 2..14 of k, s = word + bits(u): byte in the top 8 bits; certificate s * wmul_c > amb_t (mod 2^32)."""
    c = np.float32(1.0 - 1.0 / 8192.0)
    c1 = np.float32(32764.0 / 8388608.0)
    u = _f32_fma(v, np.full_like(v, c), np.full_like(v, np.float32(one_c)))
    k = _f32_fma(v, np.full_like(v, c1), np.ones_like(v), down=True)
    idx = (k.view(np.uint32) & np.uint32(0x7ffc)) >> np.uint32(2)
    s = (g8a[idx].astype(np.uint64) + u.view(np.uint32).astype(np.uint64)) & np.uint64(0xffffffff)
    byte = (s >> np.uint64(24)).astype(np.int64)
    dist = (s * np.uint64(wmul_c)) & np.uint64(0xffffffff)
    return byte, dist > np.uint64(amb_t)


@pytest.mark.parametrize("delta_override", [0.0, 1e-6, 7.9e-5])
def test_gamma_stage_and_certificate_on_the_cpu(ip, delta_override):
    """For values all over [0, 1] — a dense sweep, every threshold's neighbourhood float by float, and the ends — the
    emulated cheap stage gives the byte of the exact step function (number of thresholds <= v) wherever its certificate
    holds, and a certified value keeps that byte over the whole interval [v - delta_c, v + delta_c]: what makes a
    certified pixel's byte the reference's.  No GPU: table and constants come from ipb_spec_tables."""
    g8a, thr, one, wmul, amb_t, delta = _spec_tables(ip, delta_override)
    assert np.all(np.diff(thr) > 0) and thr[0] > 0 and thr[-1] <= 1
    sweep = np.linspace(2.0 ** -14, 1.0, 2_000_001).astype(np.float32)
    near = []
    for t in thr:   # +-4000 floats around every threshold (a few 1e-4 relative: far beyond delta near the dark end)
        b = np.array([t], np.float32).view(np.uint32)[0]
        near.append(np.arange(int(b) - 4000, int(b) + 4000, dtype=np.int64).astype(np.uint32).view(np.float32))
    v = np.unique(np.concatenate([sweep] + near + [np.array([1.0, thr[0], thr[-1]], np.float32)]))
    v = v[(v >= 2.0 ** -14) & (v <= 1.0)]
    exact = np.searchsorted(thr, v, side="right")            # number of thresholds <= v
    for ch in range(3):
        dc = np.float64(delta[1 + ch])
        byte, certified = _cheap_bytes(v, g8a, one[ch], wmul[ch], amb_t)
        assert certified.mean() > 0.5, "the certificate refuses most values"
        assert np.array_equal(byte[certified], exact[certified]), f"channel {ch}: a certified byte differs from the step function"
        lo = np.searchsorted(thr, (v.astype(np.float64) - dc).astype(np.float32), side="right")
        hi = np.searchsorted(thr, (v.astype(np.float64) + dc).astype(np.float32), side="right")
        bad = certified & ((lo != exact) | (hi != exact))
        assert not bad.any(), f"channel {ch}: {int(bad.sum())} certified values lie within delta of a threshold"
        # and the certificate is not vacuous: a value within half of delta of a threshold is refused
        close = np.min(np.abs(v[:, None].astype(np.float64) - thr[None, ::16].astype(np.float64)), axis=1) < dc / 2
        assert not (certified & close).any()
