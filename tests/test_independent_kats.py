"""Known answers that route through NEITHER copy of the host arithmetic (oracle/oracle.c and ipb_host.cu share their
statement of scaling.rs:8-23 and curves.rs:68-124): sizes worked out by hand from the reference's source, and
Fritsch-Carlson coefficients in exact rational arithmetic.  Both the oracle and the product are held to them."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

# calculate_scaling_total(width, height, maxwidth, maxheight) -> (scale, width, height), src/scaling.rs:8-23
SCALING = [
    # no limits / limits larger than the image: never up-scales (:9-10, :15-16)
    ((6000, 4000, 0, 0), (1.0, 6000, 4000)),
    ((100, 100, 200, 200), (1.0, 100, 100)),
    ((100, 100, 100, 0), (1.0, 100, 100)),
    # maxwidth only: yscale = 1.0 <= xscale = 4 -> (:19-21) width = maxwidth, height = trunc(4000 / 4)
    ((6000, 4000, 1500, 0), (4.0, 1500, 1000)),
    # maxheight only: xscale = 1.0 < yscale = 4 -> (:17-18) width = trunc(6000 / 4), height = maxheight
    ((6000, 4000, 0, 1000), (4.0, 1500, 1000)),
    # truncation, width limit: xscale = 6000/1700 = 3.529..., height = trunc(4000 * 1700 / 6000 = 1133.33) = 1133
    ((6000, 4000, 1700, 0), (6000 / 1700, 1700, 1133)),
    # yscale > xscale with truncation: yscale = 3000/700 = 4.2857, width = trunc(1000 * 700 / 3000 = 233.33) = 233
    ((1000, 3000, 600, 700), (3000 / 700, 233, 700)),
    # xscale > yscale: xscale = 5, height = trunc(3000 / 5) = 600 (the height limit of 700 is not reached)
    ((5000, 3000, 1000, 700), (5.0, 1000, 600)),
    # one axis would up-scale, the other shrinks: xscale = 0.5, yscale = 4 -> yscale wins, width = trunc(100 / 4) = 25
    ((100, 400, 200, 100), (4.0, 25, 100)),
    # equal scales take the else branch (:19): width = maxwidth, height = trunc(2000 / 4)
    ((4000, 2000, 1000, 500), (4.0, 1000, 500)),
]


@pytest.mark.parametrize("args,want", SCALING)
def test_scaling_total_hand_computed(orc, args, want):
    import imagepipe_b200 as ip
    for name, size_fn, scale_fn in (("oracle", orc.lib().orc_scaling_size, orc.lib().orc_calculate_scale),
                                    ("product", ip.lib().ipb_scaling_size, ip.lib().ipb_calculate_scale)):
        w, h = C.c_size_t(), C.c_size_t()
        size_fn(*args, C.byref(w), C.byref(h))
        assert (w.value, h.value) == want[1:], name
        assert scale_fn(*args) == np.float32(np.float32(want[0])), name


def fritsch_carlson_exact(points):
    """SplineFunc::new (curves.rs:68-124) in exact rational arithmetic: (xs, ys, c1, c2, c3)."""
    pts = [(Fraction(0), Fraction(0))] + [(Fraction(x).limit_denominator(10**6), Fraction(y).limit_denominator(10**6)) for x, y in points] \
        + [(Fraction(1), Fraction(1))]
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    dxs = [xs[i + 1] - xs[i] for i in range(len(pts) - 1)]
    ms = [(ys[i + 1] - ys[i]) / dxs[i] for i in range(len(pts) - 1)]
    c1 = [ms[0]]
    for i in range(len(dxs) - 1):
        if ms[i] * ms[i + 1] <= 0:
            c1.append(Fraction(0))
        else:
            common = dxs[i] + dxs[i + 1]
            c1.append(3 * common / ((common + dxs[i + 1]) / ms[i] + (common + dxs[i]) / ms[i + 1]))
    c1.append(ms[-1])
    c2, c3 = [], []
    for i in range(len(c1) - 1):
        inv = 1 / dxs[i]
        common = c1[i] + c1[i + 1] - 2 * ms[i]
        c2.append((ms[i] - c1[i] - common) * inv)
        c3.append(common * inv * inv)
    return xs, ys, c1, c2, c3


def spline_exact(points, v):
    xs, ys, c1, c2, c3 = fritsch_carlson_exact(points)
    v = Fraction(float(v))
    if v >= xs[-1]:
        return ys[-1]
    if v <= xs[0]:
        return ys[0]
    i = max(k for k in range(len(xs) - 1) if xs[k] <= v)
    d = v - xs[i]
    return ys[i] + c1[i] * d + c2[i] * d * d + c3[i] * d * d * d


CURVES = [[(0.5, 0.6)], [(0.25, 0.4)], [(0.3, 0.2), (0.7, 0.9)], [(0.2, 0.6), (0.5, 0.3), (0.8, 0.7)]]


def test_default_curve_coefficients_by_hand():
    """(0,0), (0.5,0.6), (1,1): slopes 1.2, 0.8; c1[1] = 3*1 / (1.5/1.2 + 1.5/0.8) = 0.96; segment 0: common = 1.2 + 0.96
    - 2.4 = -0.24, c2 = (1.2 - 1.2 + 0.24) * 2 = 0.48, c3 = -0.24 * 4 = -0.96; segment 1: common = 0.96 + 0.8 - 1.6 = 0.16,
    c2 = (0.8 - 0.96 - 0.16) * 2 = -0.64, c3 = 0.16 * 4 = 0.64."""
    _, _, c1, c2, c3 = fritsch_carlson_exact([(0.5, 0.6)])
    assert [c1, c2, c3] == [[Fraction(6, 5), Fraction(24, 25), Fraction(4, 5)], [Fraction(12, 25), Fraction(-16, 25)],
                            [Fraction(-24, 25), Fraction(16, 25)]]


@pytest.mark.parametrize("points", CURVES)
def test_oracle_spline_coefficients_against_exact_arithmetic(orc, points):
    s = orc.Spline()
    arr = (C.c_float * (2 * len(points)))(*[c for p in points for c in p])
    orc.lib().orc_spline_new(C.byref(s), arr, len(points))
    xs, ys, c1, c2, c3 = fritsch_carlson_exact(points)
    assert s.n == len(xs)
    for name, got, want in (("c1", s.c1, c1), ("c2", s.c2, c2), ("c3", s.c3, c3)):
        for i, w in enumerate(want):
            assert abs(got[i] - float(w)) <= 4e-6 * max(1.0, abs(float(w))), (name, i, got[i], float(w))
    for v in np.linspace(-0.1, 1.1, 241, dtype=np.float32):
        got = orc.lib().orc_spline_interpolate(C.byref(s), float(v))
        assert abs(got - float(spline_exact(points, v))) <= 2e-6, (points, v)


@pytest.mark.gpu
@pytest.mark.parametrize("points", CURVES)
def test_product_spline_against_exact_arithmetic(ip, ctx, points):
    v = np.linspace(-0.1, 1.1, 241, dtype=np.float32)
    got = ip.SplineFunc(points, ctx=ctx).interpolate(v)
    want = np.array([float(spline_exact(points, x)) for x in v])
    assert np.max(np.abs(got - want)) <= 2e-6
