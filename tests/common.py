"""Shared inputs for the parity tests: SURVEY.md §8d synthetic frames and literal camera metadata."""
import ctypes as C

import numpy as np

SEED = 0x1A6E51DE
XTRANS = "GBGGRGRGRBGBGBGGRGGRGGBGBGBRGRGRGGBG"  # canonical 6x6 X-Trans layout

# a fixed cam_to_xyz_normalized with negative off-diagonals (rows sum to the D65 white point)
CAM_TO_XYZ = np.array([[0.6097, 0.2053, 0.1355, 0.0],
                       [0.2762, 0.8149, -0.0911, 0.0],
                       [0.0297, -0.1206, 1.1797, 0.0]], np.float32)
WB = [2.0, 1.0, 1.5, float("nan")]


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synth_cfa(width, height, seed=SEED, row0=0):
    """v(i) = splitmix64(seed ^ i) mod 16384 — identical to the device generator ipb_synth_cfa_u16."""
    with np.errstate(over="ignore"):
        i = np.arange(row0 * width, (row0 + height) * width, dtype=np.uint64)
        v = splitmix64(np.uint64(seed) ^ i) & np.uint64(16383)
    return v.astype(np.uint16).reshape(height, width)


def smooth_cfa(width, height, seed=1):
    """A natural-looking frame (gradients + mild noise) that keeps most pixels on the LUT branch."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    base = 600 + 14000 * (0.5 + 0.5 * np.sin(x / 37.0) * np.cos(y / 23.0)) * (x + y + 1) / (width + height)
    v = base + rng.normal(0, 60, (height, width))
    return np.clip(v, 0, 16383).astype(np.uint16)


def raw_params(cfa="RGGB", crops=(0, 0, 0, 0), matrix=CAM_TO_XYZ, wb=WB, black=512.0, white=16383.0,
               points=((0.5, 0.6),), exposure=0.0, rotation=0, fliph=False, flipv=False):
    """Parameter dict understood by oracle.fill_ops and fill_ipb_ops (same literals on both sides)."""
    return {
        "gofloat": {"crop_top": crops[0], "crop_right": crops[1], "crop_bottom": crops[2], "crop_left": crops[3],
                    "is_cfa": True, "blacklevels": [black] * 4, "whitelevels": [white] * 4},
        "demosaic": {"cfa": cfa},
        "tolab": {"cam_to_xyz": matrix, "cam_to_xyz_normalized": matrix, "wb_coeffs": list(wb)},
        "basecurve": {"exposure": exposure, "points": list(points)},
        "transform": {"rotation": rotation, "fliph": fliph, "flipv": flipv},
    }


def fill_ipb_ops(ops, params):
    """Fill an imagepipe_b200.PipelineOps from the params dict (mirror of oracle.fill_ops)."""
    g = params.get("gofloat", {})
    for k in ("crop_top", "crop_right", "crop_bottom", "crop_left"):
        setattr(ops.gofloat, k, g.get(k, 0))
    if "is_cfa" in g:
        ops.gofloat.is_cfa = int(g["is_cfa"])
    for i in range(4):
        if "blacklevels" in g:
            ops.gofloat.blacklevels[i] = g["blacklevels"][i]
        if "whitelevels" in g:
            ops.gofloat.whitelevels[i] = g["whitelevels"][i]
    if "demosaic" in params:
        ops.demosaic.cfa = params["demosaic"]["cfa"].encode()
    r = params.get("rotatecrop", {})
    for k in ("crop_top", "crop_right", "crop_bottom", "crop_left", "rotation"):
        if k in r:
            setattr(ops.rotatecrop, k, r[k])
    t = params.get("tolab", {})
    for name, rows, cols in (("cam_to_xyz", 3, 4), ("cam_to_xyz_normalized", 3, 4), ("xyz_to_cam", 4, 3)):
        if name in t:
            m = np.asarray(t[name], np.float32).reshape(rows, cols)
            for i in range(rows):
                for j in range(cols):
                    getattr(ops.tolab, name)[i][j] = m[i, j]
    if "wb_coeffs" in t:
        for i in range(4):
            ops.tolab.wb_coeffs[i] = t["wb_coeffs"][i]
    b = params.get("basecurve")
    if b is not None:
        ops.basecurve.exposure = b.get("exposure", 0.0)
        ops.basecurve.set_points(b.get("points", []))
    tr = params.get("transform", {})
    ops.transform.rotation = tr.get("rotation", 0)
    ops.transform.fliph = int(tr.get("fliph", False))
    ops.transform.flipv = int(tr.get("flipv", False))


def make_ipb_pipeline(ip, data, kind="raw", params=None, settings=None, ctx=None, on_device=False):
    if kind == "raw":
        src = ip.ImageSource.Raw(data)
    else:
        src = ip.ImageSource.Other(data)
    if on_device:
        d = ip.DeviceArray.from_numpy(data, ctx)
        src = ip.ImageSource(src.kind, src.width, src.height, src.cpp, d)
    p = ip.Pipeline.new_from_source(src, ctx=ctx)
    if params:
        fill_ipb_ops(p.ops, params)
    for k, v in (settings or {}).items():
        setattr(p.globals.settings, k, int(v))
    return p


def bits(a):
    """Bit pattern view for exact f32 comparison (NaNs with equal payload compare equal)."""
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_exact(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if got.dtype == np.float32:
        g, w = bits(got), bits(want)
        # +0.0 and -0.0 are distinct bit patterns.  Any NaN equals any NaN: x86 and the GPU generate different
        # default-NaN payloads for the same invalid operation, and the reference does not define them.
        bad = (g != w) & ~(np.isnan(got) & np.isnan(want))
    else:
        bad = got != want
    n = int(bad.sum())
    if n:
        idx = np.argwhere(bad)[:5]
        detail = "; ".join(f"{tuple(i)}: got {got[tuple(i)]!r} want {want[tuple(i)]!r}" for i in idx)
        raise AssertionError(f"{what}: {n} of {bad.size} elements differ ({detail})")
