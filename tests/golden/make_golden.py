#!/usr/bin/env python
"""Golden vectors for the parts of the path the reference's own tests do not pin (gofloat::run_raw,
demosaic::full, scaled_demosaic, to_lab with a camera matrix, basecurve with an interior point).

The reference is Rust and cannot run here, so these vectors come from a SECOND, independent restatement: scalar
numpy-float32 loops written directly from the reference source (file:line cited at each function), sharing no code
with oracle/oracle.c or the CUDA kernels.  Two independent restatements agreeing bit for bit is the strongest pin
available without the reference binary.  Frames are tiny (hand-checkable: `full_rggb_10x10` has a worked example in
tests/test_golden.py).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

Every case is bit-exact: gofloat, demosaic::full and scaled_demosaic are pure f32 arithmetic; the colour chain's
8193-entry tables call glibc's cbrtf/powf through ctypes — the libm Rust's f32::cbrt / f32::powf resolve to on
Linux — so they equal the oracle's and the product's host-built tables bit for bit.
"""
import ctypes
import os

import numpy as np

F = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))
XTRANS = "GBGGRGRGRBGBGBGGRGGRGGBGBGBRGRGRGGBG"
CAM_TO_XYZ = np.array([[0.6097, 0.2053, 0.1355, 0.0],
                       [0.2762, 0.8149, -0.0911, 0.0],
                       [0.0297, -0.1206, 1.1797, 0.0]], F)


def color_at(pattern, row, col):
    """rawloader::CFA::color_at: pattern string tiled over the sensor (2x2 / 6x6 / 2x8 / 12x12), R,G,B,E -> 0..3."""
    w, h = {4: (2, 2), 36: (6, 6), 16: (2, 8), 144: (12, 12)}[len(pattern)]
    return {"R": 0, "G": 1, "B": 2, "E": 3, "M": 1, "Y": 3}[pattern[(row % h) * w + (col % w)]]


def gofloat_cfa(raw, black, white, crops):
    """gofloat.rs:74-82 (size_image) + :122-130: ((v as f32 - mins[0]) / ranges[0]).min(1.0) on the cropped area."""
    top, right, bottom, left = crops
    oh, ow = raw.shape
    x, y, w, h = left, top, ow - left - right, oh - top - bottom
    rng = F(white) - F(black)
    out = np.zeros((h, w), F)
    for r in range(h):
        for c in range(w):
            v = (F(raw[r + y, c + x]) - F(black)) / rng
            out[r, c] = v if v < F(1.0) else F(1.0)
    return out


def demosaic_full(a, pattern):
    """demosaic.rs:67-119: per-colour mean of the 3x3 neighbourhood; same-colour neighbours are discarded, the centre is
    kept, out-of-frame taps do not count."""
    h, w = a.shape
    out = np.zeros((h, w, 4), F)
    for row in range(h):
        for col in range(w):
            pix = color_at(pattern, row, col)
            sums, counts = [F(0)] * 5, [F(0)] * 5
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    oc = color_at(pattern, row + 48 + dy, col + 48 + dx)
                    b = oc if (oc != pix or (dx == 0 and dy == 0)) else 4
                    r2, c2 = row + dy, col + dx
                    if 0 <= r2 < h and 0 <= c2 < w:
                        sums[b] = F(sums[b] + a[r2, c2])
                        counts[b] = F(counts[b] + F(1))
            for c in range(4):
                if counts[c] > 0:
                    out[row, col, c] = F(sums[c] / counts[c])
    return out


def scaled_demosaic(a, pattern, nw, nh):
    """scaling.rs:132-136 -> :35-48 -> transform_buffer :51-130 with corners (0,0), (w-1,0), (0,h-1), CFA mode."""
    h, w = a.shape
    sxx = F(F(w - 1) - F(0)) / F(nw - 1)
    sxy = F(F(0) - F(0)) / F(nw - 1)
    syx = F(F(0) - F(0)) / F(nh - 1)
    syy = F(F(h - 1) - F(0)) / F(nh - 1)
    out = np.zeros((nh, nw, 4), F)
    fl = lambda v: int(np.floor(v))
    for row in range(nh):
        rfx = F(F(0) + F(syx * F(row)))
        rtx = F(F(0) + F(syx * F(row + 1)))
        rfy = F(F(0) + F(syy * F(row)))
        rty = F(F(0) + F(syy * F(row + 1)))
        rcx = F(F(F(F(0) + F(syx * F(row))) + F(syx / F(2))) - F(0.5))
        rcy = F(F(F(F(0) + F(syy * F(row))) + F(syy / F(2))) - F(0.5))
        for col in range(nw):
            from_x = min(w - 1, fl(F(rfx + F(sxx * F(col)))))
            to_x = min(w - 1, fl(F(rtx + F(sxx * F(col + 1)))))
            from_y = min(h - 1, fl(F(rfy + F(sxy * F(col)))))
            to_y = min(h - 1, fl(F(rty + F(sxy * F(col + 1)))))
            cx = F(F(rcx + F(sxx * F(col))) + F(sxx / F(2)))
            cy = F(F(rcy + F(sxy * F(col))) + F(sxy / F(2)))
            sums, counts = [F(0)] * 4, [F(0)] * 4
            for y in range(from_y, to_y + 1):
                for x in range(from_x, to_x + 1):
                    dx = F(F(F(x) - cx) / sxx)
                    dy = F(F(F(y) - cy) / syy)
                    f = F(F(F(1) - F(dx * dx)) - F(dy * dy))
                    if f < 0:
                        f = F(0)
                    c = color_at(pattern, y, x)
                    sums[c] = F(sums[c] + F(a[y, x] * f))
                    counts[c] = F(counts[c] + f)
            for c in range(4):
                if counts[c] > 0:
                    out[row, col, c] = F(sums[c] / counts[c])
    return out


# ---- colour chain (color_conversions.rs, colorspaces.rs, curves.rs, gamma.rs), numpy f32 ------------------------------

E_ = F(216.0) / F(24389.0)
K_ = F(24389.0) / F(27.0)


def table(fn):
    """TransformLookup::new (color_conversions.rs:86-100): 8193 entries fn(i / 8191)."""
    return np.array([fn(F(F(i) / F(8191.0))) for i in range(8193)], F)


def lookup(t, fn, v):
    """TransformLookup::lookup (:102-114)."""
    if v < 0 or v > 1:
        return fn(v)
    pos = F(v * F(8191.0))
    key = int(pos)
    a = F(pos - F(np.trunc(pos)))
    return F(t[key] + F(a * F(t[key + 1] - t[key])))


_libm = ctypes.CDLL("libm.so.6")   # Rust's f32::cbrt / f32::powf call the platform libm: use the same one
_libm.cbrtf.restype = _libm.powf.restype = ctypes.c_float
_libm.cbrtf.argtypes = [ctypes.c_float]
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]


def f_lab(v):
    """color_conversions.rs:120-124"""
    return F(_libm.cbrtf(float(v))) if v > E_ else F(F(F(K_ * v) + F(16.0)) / F(116.0))


def f_gamma(v):
    """color_conversions.rs:134-140"""
    if v < F(0.0031308):
        return F(v * F(12.92))
    return F(F(F(1.055) * F(_libm.powf(float(v), float(F(F(1.0) / F(2.4)))))) - F(0.055))


def inverse33(m):
    """color_conversions.rs:20-39: adjugate times 1/det, every step rounded to f32."""
    def d(a, b, c, e):  # a*b - c*e
        return F(F(a * b) - F(c * e))
    det = F(F(F(m[0][0] * d(m[1][1], m[2][2], m[2][1], m[1][2])) - F(m[0][1] * d(m[1][0], m[2][2], m[1][2], m[2][0])))
            + F(m[0][2] * d(m[1][0], m[2][1], m[1][1], m[2][0])))
    inv = F(F(1.0) / det)
    o = np.zeros((3, 3), F)
    o[0][0] = F(d(m[1][1], m[2][2], m[2][1], m[1][2]) * inv)
    o[0][1] = F(-d(m[0][1], m[2][2], m[0][2], m[2][1]) * inv)
    o[0][2] = F(d(m[0][1], m[1][2], m[0][2], m[1][1]) * inv)
    o[1][0] = F(-d(m[1][0], m[2][2], m[1][2], m[2][0]) * inv)
    o[1][1] = F(d(m[0][0], m[2][2], m[0][2], m[2][0]) * inv)
    o[1][2] = F(-d(m[0][0], m[1][2], m[1][0], m[0][2]) * inv)
    o[2][0] = F(d(m[1][0], m[2][1], m[2][0], m[1][1]) * inv)
    o[2][1] = F(-d(m[0][0], m[2][1], m[2][0], m[0][1]) * inv)
    o[2][2] = F(d(m[0][0], m[1][1], m[1][0], m[0][1]) * inv)
    return o


SRGB_D65_33 = np.array([[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750],
                        [0.0193339, 0.1191920, 0.9503041]], F)


def colour_chain(rgbe, mul, cmatrix, points, lab_t, gam_t):
    """to_lab (colorspaces.rs:89-112, color_conversions.rs:42-55,156-169), basecurve (curves.rs:33-157),
    from_lab (:58-65,172-191), gamma (gamma.rs:16-26) for one pixel."""
    c = [min(F(rgbe[i] * mul[i]), F(1.0)) for i in range(4)]
    xyz = [F(F(F(F(c[0] * cmatrix[r][0]) + F(c[1] * cmatrix[r][1])) + F(c[2] * cmatrix[r][2])) + F(c[3] * cmatrix[r][3]))
           for r in range(3)]
    xr, yr, zr = F(xyz[0] / F(0.95047)), F(xyz[1] / F(1.0)), F(xyz[2] / F(1.08883))
    fx, fy, fz = (lookup(lab_t, f_lab, v) for v in (xr, yr, zr))
    L = F(F(F(116.0) * fy) - F(16.0))
    A = F(F(500.0) * F(fx - fy))
    B = F(F(200.0) * F(fy - fz))
    lab = [F(L / F(100.0)), F(F(A + F(127.0)) / F(255.0)), F(F(B + F(127.0)) / F(255.0))]
    lab[0] = spline(points, lab[0])
    cl, ca, cb = F(lab[0] * F(100.0)), F(F(lab[1] * F(255.0)) - F(127.0)), F(F(lab[2] * F(255.0)) - F(127.0))
    fy = F(F(cl + F(16.0)) / F(116.0))
    fx = F(F(ca / F(500.0)) + fy)
    fz = F(fy - F(cb / F(200.0)))
    fx3, fz3 = F(F(fx * fx) * fx), F(F(fz * fz) * fz)
    xr = fx3 if fx3 > E_ else F(F(F(F(116.0) * fx) - F(16.0)) / K_)
    yr = F(F(fy * fy) * fy) if cl > F(K_ * E_) else F(cl / K_)
    zr = fz3 if fz3 > E_ else F(F(F(F(116.0) * fz) - F(16.0)) / K_)
    X, Y, Z = F(xr * F(0.95047)), F(yr * F(1.0)), F(zr * F(1.08883))
    m = inverse33(SRGB_D65_33)
    rgb = [F(F(F(X * m[r][0]) + F(Y * m[r][1])) + F(Z * m[r][2])) for r in range(3)]
    return [lookup(gam_t, f_gamma, min(max(v, F(0.0)), F(1.0))) for v in rgb], lab


def spline(points, v):
    """SplineFunc::new + interpolate (curves.rs:68-157): monotone cubic (Fritsch-Carlson) through
    (0,0), points..., (1,1)."""
    pts = [(F(0), F(0))] + [(F(x), F(y)) for x, y in points] + [(F(1), F(1))]
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    n = len(pts)
    dxs = [F(xs[i + 1] - xs[i]) for i in range(n - 1)]
    ms = [F(F(ys[i + 1] - ys[i]) / dxs[i]) for i in range(n - 1)]
    c1 = [ms[0]]
    for i in range(n - 2):
        m, mn = ms[i], ms[i + 1]
        if m * mn <= 0:
            c1.append(F(0))
        else:
            dx, dxn = dxs[i], dxs[i + 1]
            common = F(dx + dxn)
            c1.append(F(F(F(3.0) * common) / F(F(F(common + dxn) / m) + F(F(common + dx) / mn))))
    c1.append(ms[-1])
    c2, c3 = [], []
    for i in range(n - 1):
        c1v, m = c1[i], ms[i]
        inv = F(F(1.0) / dxs[i])
        common = F(F(F(c1v + c1[i + 1]) - m) - m)
        c2.append(F(F(F(m - c1v) - common) * inv))
        c3.append(F(F(common * inv) * inv))
    if v >= xs[-1]:
        return ys[-1]
    if v <= xs[0]:
        return ys[0]
    i = max(k for k in range(n - 1) if xs[k] <= v)
    d = F(v - xs[i])
    return F(F(F(ys[i] + F(c1[i] * d)) + F(F(c2[i] * d) * d)) + F(F(F(c3[i] * d) * d) * d))


def synth(shape, seed, lo=0, hi=1024):
    return np.random.default_rng(seed).integers(lo, hi, shape).astype(np.uint16)


def main():
    cases = {}
    raw = synth((10, 10), 1)
    raw[0, 0], raw[2, 3], raw[9, 9] = 10, 1023, 700       # below black, at white, ordinary
    g = gofloat_cfa(raw, 64.0, 1023.0, (0, 0, 0, 0))
    cases["full_rggb_10x10"] = dict(raw=raw, black=64.0, white=1023.0, crops=(0, 0, 0, 0), cfa="RGGB", gofloat=g,
                                  demosaic=demosaic_full(g, "RGGB"))
    raw = synth((14, 16), 2)
    g = gofloat_cfa(raw, 60.0, 1000.0, (1, 2, 1, 2))      # crops: 12x12 left
    cases["full_xtrans_12x12"] = dict(raw=raw, black=60.0, white=1000.0, crops=(1, 2, 1, 2), cfa=XTRANS, gofloat=g,
                                      demosaic=demosaic_full(g, XTRANS))
    raw = synth((11, 13), 3)
    g = gofloat_cfa(raw, 0.0, 1023.0, (0, 0, 0, 0))
    cases["full_gbrg_11x13"] = dict(raw=raw, black=0.0, white=1023.0, crops=(0, 0, 0, 0), cfa="GBRG", gofloat=g,
                                    demosaic=demosaic_full(g, "GBRG"))
    raw = synth((16, 16), 4)
    g = gofloat_cfa(raw, 32.0, 1023.0, (0, 0, 0, 0))
    cases["scaled_rggb_16x16_to_4x4"] = dict(raw=raw, black=32.0, white=1023.0, crops=(0, 0, 0, 0), cfa="RGGB",
                                             gofloat=g, nwidth=4, nheight=4, demosaic=scaled_demosaic(g, "RGGB", 4, 4))
    raw = synth((24, 30), 5)
    g = gofloat_cfa(raw, 32.0, 1023.0, (0, 0, 0, 0))
    cases["scaled_xtrans_30x24_to_10x8"] = dict(raw=raw, black=32.0, white=1023.0, crops=(0, 0, 0, 0), cfa=XTRANS,
                                                gofloat=g, nwidth=10, nheight=8,
                                                demosaic=scaled_demosaic(g, XTRANS, 10, 8))
    # colour chain on 96 RGBE pixels (in range, clipped, negative, saturated)
    rng = np.random.default_rng(6)
    px = rng.uniform(-0.03, 1.1, (96, 4)).astype(F)
    px[:, 3] = 0
    px[0] = [0, 0, 0, 0]
    px[1] = [1, 1, 1, 0]
    px[2] = [0.18, 0.18, 0.18, 0]
    wb = [2.0, 1.0, 1.5, float("nan")]
    mul = [F(v / wb[1]) if np.isfinite(v) and v != 0 else F(1.0) for v in wb]   # normalize_wbs, colorspaces.rs:12-27
    lab_t, gam_t = table(f_lab), table(f_gamma)
    pts = [(0.5, 0.6)]
    rgb, lab = zip(*(colour_chain(p, mul, CAM_TO_XYZ, pts, lab_t, gam_t) for p in px))
    cases["colour_chain_96px"] = dict(rgbe=px, wb=np.array(wb, F), matrix=CAM_TO_XYZ, points=np.array(pts, F),
                                      lab=np.array(lab, F), rgb=np.array(rgb, F))
    for name, d in cases.items():
        np.savez(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in d.items()})
        print("wrote", name, {k: np.asarray(v).shape for k, v in d.items()})


if __name__ == "__main__":
    main()
