"""ctypes binding of the CPU parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (imagepipe_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX_CURVE_POINTS = 32
SRC_RAW_U16, SRC_RAW_F32, SRC_RGB8, SRC_RGB16 = 0, 1, 2, 3
ROT_NORMAL, ROT_90, ROT_180, ROT_270 = 0, 1, 2, 3


def build(force=False):
    """Compile oracle/liboracle.so from oracle/*.c (gcc, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "kats.c", "lanczos.c", "oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


NATIVE_FLAGS = "-O3 -march=native -ffp-contract=off -fno-fast-math -fopenmp"


def use_native():
    """bench.py's CPU-baseline legs only: rebuild the same sources with BASELINE.md's flags (-O3 -march=native) on THIS
    machine and route lib() to that build.  Returns the flags actually in use (the portable build's when gcc is absent)."""
    global _LIB_PATH, _lib
    path = os.path.join(_HERE, "liboracle_native.so")
    try:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle_native.so"], check=True, capture_output=True)
    except Exception:
        return "-O2 -march=x86-64-v2 -ffp-contract=off -fopenmp (portable checker build; native rebuild failed)"
    _LIB_PATH, _lib = path, None
    return NATIVE_FLAGS


class Buffer(C.Structure):
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("colors", C.c_size_t),
                ("monochrome", C.c_int), ("data", C.POINTER(C.c_float))]


class GoFloat(C.Structure):
    _fields_ = [("crop_top", C.c_size_t), ("crop_right", C.c_size_t), ("crop_bottom", C.c_size_t),
                ("crop_left", C.c_size_t), ("is_cfa", C.c_int),
                ("blacklevels", C.c_float * 4), ("whitelevels", C.c_float * 4)]


class Demosaic(C.Structure):
    _fields_ = [("cfa", C.c_char * 148)]


class RotateCrop(C.Structure):
    _fields_ = [("crop_top", C.c_float), ("crop_right", C.c_float), ("crop_bottom", C.c_float),
                ("crop_left", C.c_float), ("rotation", C.c_float), ("input_ratio", C.c_float),
                ("has_output_size", C.c_int), ("output_width", C.c_size_t), ("output_height", C.c_size_t)]


class ToLab(C.Structure):
    _fields_ = [("cam_to_xyz", (C.c_float * 4) * 3), ("cam_to_xyz_normalized", (C.c_float * 4) * 3),
                ("xyz_to_cam", (C.c_float * 3) * 4), ("wb_coeffs", C.c_float * 4)]


class BaseCurve(C.Structure):
    _fields_ = [("exposure", C.c_float), ("npoints", C.c_size_t),
                ("points", (C.c_float * 2) * MAX_CURVE_POINTS)]


class Transform(C.Structure):
    _fields_ = [("rotation", C.c_int), ("fliph", C.c_int), ("flipv", C.c_int)]


class Settings(C.Structure):
    _fields_ = [("maxwidth", C.c_size_t), ("maxheight", C.c_size_t), ("demosaic_width", C.c_size_t),
                ("demosaic_height", C.c_size_t), ("linear", C.c_int), ("use_fastpath", C.c_int)]


class Source(C.Structure):
    _fields_ = [("kind", C.c_int), ("width", C.c_size_t), ("height", C.c_size_t), ("cpp", C.c_size_t),
                ("data", C.c_void_p)]


class Ops(C.Structure):
    _fields_ = [("gofloat", GoFloat), ("demosaic", Demosaic), ("rotatecrop", RotateCrop), ("tolab", ToLab),
                ("basecurve", BaseCurve), ("transform", Transform)]


class Pipeline(C.Structure):
    _fields_ = [("image", Source), ("settings", Settings), ("ops", Ops)]


class Cfa(C.Structure):
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("pattern", (C.c_uint8 * 48) * 48)]


class Spline(C.Structure):
    _N = MAX_CURVE_POINTS + 2
    _fields_ = [("n", C.c_size_t), ("x", C.c_float * _N), ("y", C.c_float * _N), ("c1", C.c_float * _N),
                ("c2", C.c_float * _N), ("c3", C.c_float * _N), ("nseg", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    BP = C.POINTER(Buffer)
    sz = C.c_size_t
    szp = C.POINTER(C.c_size_t)
    fp = C.POINTER(C.c_float)
    sigs = {
        "orc_buffer_new": (BP, [sz, sz, sz, C.c_int]),
        "orc_buffer_from": (BP, [sz, sz, sz, C.c_int, fp]),
        "orc_buffer_free": (None, [BP]),
        "orc_set_threads": (None, [C.c_int]),
        "orc_get_threads": (C.c_int, []),
        "orc_matrices": (None, [fp, fp]),
        "orc_lut_xyz_lab": (fp, []),
        "orc_lut_srgb_reverse": (fp, []),
        "orc_lut_srgb_transform": (fp, []),
        "orc_expand_srgb_gamma": (C.c_float, [C.c_float]),
        "orc_apply_srgb_gamma": (C.c_float, [C.c_float]),
        "orc_xyz_to_lab": (None, [C.c_float, C.c_float, C.c_float, fp]),
        "orc_lab_to_xyz": (None, [C.c_float, C.c_float, C.c_float, fp]),
        "orc_camera_to_lab": (None, [fp, fp, fp, fp]),
        "orc_lab_to_rgb": (None, [fp, fp, fp]),
        "orc_input8bit": (C.c_float, [C.c_uint8]),
        "orc_input16bit": (C.c_float, [C.c_uint16]),
        "orc_output8bit": (C.c_uint8, [C.c_float]),
        "orc_output16bit": (C.c_uint16, [C.c_float]),
        "orc_cfa_new": (C.c_int, [C.POINTER(Cfa), C.c_char_p]),
        "orc_cfa_color_at": (sz, [C.POINTER(Cfa), sz, sz]),
        "orc_calculate_scale": (C.c_float, [sz, sz, sz, sz]),
        "orc_scaling_size": (None, [sz, sz, sz, sz, szp, szp]),
        "orc_transform_buffer_f32": (None, [fp, sz, sz, C.POINTER(C.c_long), C.POINTER(C.c_long),
                                            C.POINTER(C.c_long), sz, sz, sz, C.POINTER(Cfa), fp]),
        "orc_scaled_demosaic": (BP, [C.POINTER(Cfa), BP, sz, sz]),
        "orc_scale_down_opbuf": (BP, [BP, sz, sz]),
        "orc_lanczos_resize": (BP, [BP, sz, sz, C.c_int]),
        "orc_lanczos_weights": (None, [sz, sz, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), fp, C.POINTER(C.c_size_t)]),
        "orc_scale_down_srgb": (None, [C.c_void_p, sz, sz, sz, sz, C.c_void_p]),
        "orc_scale_down_srgb16": (None, [C.c_void_p, sz, sz, sz, sz, C.c_void_p]),
        "orc_spline_new": (None, [C.POINTER(Spline), C.c_void_p, sz]),
        "orc_spline_interpolate": (C.c_float, [C.POINTER(Spline), C.c_float]),
        "orc_gofloat_size_image": (None, [C.POINTER(GoFloat), sz, sz, szp]),
        "orc_gofloat_run": (BP, [C.POINTER(GoFloat), C.POINTER(Source)]),
        "orc_demosaic_full": (BP, [C.POINTER(Cfa), BP]),
        "orc_demosaic_run": (BP, [C.POINTER(Demosaic), C.POINTER(Settings), BP]),
        "orc_rotatecrop_run": (BP, [C.POINTER(RotateCrop), BP]),
        "orc_rotatecrop_transform_forward": (None, [C.POINTER(RotateCrop), sz, sz, szp, szp]),
        "orc_rotatecrop_transform_reverse": (None, [C.POINTER(RotateCrop), sz, sz, szp, szp]),
        "orc_rotatecrop_reset": (None, [C.POINTER(RotateCrop)]),
        "orc_tolab_run": (BP, [C.POINTER(ToLab), BP]),
        "orc_basecurve_run": (BP, [C.POINTER(BaseCurve), BP]),
        "orc_fromlab_run": (BP, [BP]),
        "orc_gamma_run": (BP, [C.POINTER(Settings), BP]),
        "orc_orientation_flips": (None, [C.POINTER(Transform), C.POINTER(C.c_int)]),
        "orc_rotate_buffer": (BP, [BP, C.c_int, C.c_int, C.c_int]),
        "orc_transform_run": (BP, [C.POINTER(Transform), BP]),
        "orc_transform_transform_forward": (None, [C.POINTER(Transform), sz, sz, szp, szp]),
        "orc_pipeline_defaults": (None, [C.POINTER(Pipeline), C.POINTER(Source)]),
        "orc_pipeline_negotiate": (None, [C.POINTER(Pipeline), szp, szp]),
        "orc_pipeline_run": (BP, [C.POINTER(Pipeline)]),
        "orc_pipeline_last_timings": (None, [C.POINTER(C.c_double)]),
        "orc_pipeline_output_8bit": (C.c_int, [C.POINTER(Pipeline), C.POINTER(C.c_void_p), szp, szp]),
        "orc_pipeline_output_16bit": (C.c_int, [C.POINTER(Pipeline), C.POINTER(C.c_void_p), szp, szp]),
        "orc_free": (None, [C.c_void_p]),
    }
    for kat in ("roundtrip_8bit", "roundtrip_16bit", "roundtrip_8bit_gamma", "roundtrip_16bit_gamma",
                "roundtrip_8bit_lab_xyz", "roundtrip_8bit_lab_rgb", "roundtrip_8bit_lab_rgb_gamma",
                "roundtrip_16bit_lab_xyz", "roundtrip_16bit_lab_rgb", "roundtrip_16bit_lab_rgb_gamma",
                "rotatecrop_roundtrip_transform", "rotatecrop_roundtrip_transform_rotation"):
        sigs["orc_kat_" + kat] = (C.c_long, [])
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# ---------------------------------------------------------------------------- helpers

def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def buffer_from_numpy(arr, monochrome=False):
    """arr: float32 (h, w, colors) or (h, w) -> orc_buffer* (owned by the caller; free it)."""
    a = np.ascontiguousarray(arr, dtype=np.float32)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    return lib().orc_buffer_from(w, h, c, int(monochrome), _fptr(a))


def buffer_to_numpy(bp, free=True):
    b = bp.contents
    n = b.width * b.height * b.colors
    out = np.ctypeslib.as_array(b.data, shape=(n,)).copy().reshape(b.height, b.width, b.colors) if n else \
        np.zeros((b.height, b.width, b.colors), np.float32)
    mono = bool(b.monochrome)
    if free:
        lib().orc_buffer_free(bp)
    return out, mono


def fill_ops(ops, params):
    """Fill an Ops struct from the params dict used throughout tests/ (see tests/common.py)."""
    g = params.get("gofloat", {})
    ops.gofloat.crop_top = g.get("crop_top", 0)
    ops.gofloat.crop_right = g.get("crop_right", 0)
    ops.gofloat.crop_bottom = g.get("crop_bottom", 0)
    ops.gofloat.crop_left = g.get("crop_left", 0)
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
        pts = b.get("points", [])
        ops.basecurve.npoints = len(pts)
        for i, (x, y) in enumerate(pts):
            ops.basecurve.points[i][0] = x
            ops.basecurve.points[i][1] = y
    tr = params.get("transform", {})
    ops.transform.rotation = tr.get("rotation", 0)
    ops.transform.fliph = int(tr.get("fliph", False))
    ops.transform.flipv = int(tr.get("flipv", False))


_KINDS = {("raw", np.dtype(np.uint16)): SRC_RAW_U16, ("raw", np.dtype(np.float32)): SRC_RAW_F32,
          ("rgb", np.dtype(np.uint8)): SRC_RGB8, ("rgb", np.dtype(np.uint16)): SRC_RGB16}


def make_source(data, kind="raw", cpp=None):
    """data: numpy (h, w) / (h, w, cpp).  Returns (Source, keepalive array)."""
    a = np.ascontiguousarray(data)
    h, w = a.shape[:2]
    if cpp is None:
        cpp = a.shape[2] if a.ndim == 3 else 1
    s = Source(_KINDS[(kind, a.dtype)], w, h, cpp, a.ctypes.data)
    return s, a


def make_pipeline(data, kind="raw", params=None, settings=None, cpp=None):
    src, keep = make_source(data, kind, cpp)
    p = Pipeline()
    lib().orc_pipeline_defaults(C.byref(p), C.byref(src))
    if params:
        fill_ops(p.ops, params)
    for k, v in (settings or {}).items():
        setattr(p.settings, k, int(v))
    p._keep = keep
    return p


def pipeline_run(p):
    bp = lib().orc_pipeline_run(C.byref(p))
    if not bp:
        raise RuntimeError("oracle pipeline_run failed")
    return buffer_to_numpy(bp)[0]


def pipeline_output_8bit(p):
    d = C.c_void_p()
    w = C.c_size_t()
    h = C.c_size_t()
    rc = lib().orc_pipeline_output_8bit(C.byref(p), C.byref(d), C.byref(w), C.byref(h))
    if rc:
        raise RuntimeError("oracle output_8bit failed")
    out = np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_uint8)), shape=(h.value * w.value * 3,)).copy()
    lib().orc_free(d)
    return out.reshape(h.value, w.value, 3)


def pipeline_output_16bit(p):
    d = C.c_void_p()
    w = C.c_size_t()
    h = C.c_size_t()
    rc = lib().orc_pipeline_output_16bit(C.byref(p), C.byref(d), C.byref(w), C.byref(h))
    if rc:
        raise RuntimeError("oracle output_16bit failed")
    out = np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_uint16)), shape=(h.value * w.value * 3,)).copy()
    lib().orc_free(d)
    return out.reshape(h.value, w.value, 3)


def last_timings():
    t = (C.c_double * 8)()
    lib().orc_pipeline_last_timings(t)
    return dict(zip(("gofloat", "demosaic", "rotatecrop", "to_lab", "basecurve", "from_lab", "gamma", "transform"),
                    list(t)))
