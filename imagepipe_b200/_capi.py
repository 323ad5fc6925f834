"""ctypes binding of libipb200.so (include/ipb200.h).

This module only loads the CUDA library and declares its signatures; there is no Python or CPU
implementation of any op behind it.  If the library is missing or no B200 is visible, calls fail loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libipb200.so")

IPB_OK = 0
ERR_NAMES = {1: "IPB_ERR_INVALID", 2: "IPB_ERR_BAD_COLORS", 3: "IPB_ERR_BAD_CFA", 4: "IPB_ERR_CUDA",
             5: "IPB_ERR_UNSUPPORTED", 6: "IPB_ERR_NOMEM"}
MAX_CURVE_POINTS = 32
SRC_RAW_U16, SRC_RAW_F32, SRC_RGB8, SRC_RGB16 = 0, 1, 2, 3
ROT_NORMAL, ROT_90, ROT_180, ROT_270 = 0, 1, 2, 3


class IpbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


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
                ("data", C.c_void_p), ("on_device", C.c_int)]


class Ops(C.Structure):
    _fields_ = [("gofloat", GoFloat), ("demosaic", Demosaic), ("rotatecrop", RotateCrop), ("tolab", ToLab),
                ("basecurve", BaseCurve), ("transform", Transform)]


class Stripe(C.Structure):
    _fields_ = [("full_height", C.c_size_t), ("src_row0", C.c_size_t), ("out_row0", C.c_size_t),
                ("out_row1", C.c_size_t)]


class Halo(C.Structure):
    """ipb_halo: byte offsets / sizes of the rows exchanged with the stripe above (rank - 1) and below (rank + 1)."""
    _fields_ = [(n, C.c_size_t) for n in ("send_up_off", "send_up_bytes", "recv_up_off", "recv_up_bytes",
                                          "send_down_off", "send_down_bytes", "recv_down_off", "recv_down_bytes")]


COMM_ID_BYTES = 128

_lib = None


def lib():
    """Load libipb200.so.  Raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C imagepipe_b200/csrc` "
                          "(imagepipe_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i, f = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    szp, vpp = C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)
    P = C.POINTER
    sigs = {
        "ipb_version": (i, []),
        "ipb_ctx_create": (i, [i, vp, vpp]),
        "ipb_ctx_destroy": (None, [vp]),
        "ipb_ctx_set_stream": (i, [vp, vp]),
        "ipb_ctx_synchronize": (i, [vp]),
        "ipb_last_error": (C.c_char_p, [vp]),
        "ipb_ctx_launch_count": (C.c_ulonglong, [vp]),
        "ipb_host_alloc": (i, [sz, vpp]),
        "ipb_host_free": (None, [vp]),
        "ipb_device_alloc": (i, [vp, sz, vpp]),
        "ipb_device_free": (i, [vp, vp]),
        "ipb_device_upload": (i, [vp, vp, vp, sz]),
        "ipb_device_download": (i, [vp, vp, vp, sz]),
        "ipb_buffer_new": (i, [vp, sz, sz, sz, i, vpp]),
        "ipb_buffer_upload": (i, [vp, sz, sz, sz, i, vp, vpp]),
        "ipb_buffer_wrap": (i, [vp, sz, sz, sz, i, vp, vpp]),
        "ipb_buffer_download": (i, [vp, vp, vp]),
        "ipb_buffer_retain": (None, [vp]),
        "ipb_buffer_release": (None, [vp]),
        "ipb_buffer_width": (sz, [vp]),
        "ipb_buffer_height": (sz, [vp]),
        "ipb_buffer_colors": (sz, [vp]),
        "ipb_buffer_monochrome": (i, [vp]),
        "ipb_buffer_device_ptr": (vp, [vp]),
        "ipb_gofloat_run": (i, [vp, vp, vp, vpp]),
        "ipb_demosaic_run": (i, [vp, vp, vp, vp, vpp]),
        "ipb_rotatecrop_run": (i, [vp, vp, vp, vpp]),
        "ipb_tolab_run": (i, [vp, vp, vp, vpp]),
        "ipb_basecurve_run": (i, [vp, vp, vp, vpp]),
        "ipb_fromlab_run": (i, [vp, vp, vpp]),
        "ipb_gamma_run": (i, [vp, vp, vp, vpp]),
        "ipb_transform_run": (i, [vp, vp, vp, vpp]),
        "ipb_gofloat_transform_forward": (None, [vp, sz, sz, szp, szp]),
        "ipb_rotatecrop_transform_forward": (None, [vp, sz, sz, szp, szp]),
        "ipb_rotatecrop_transform_reverse": (None, [vp, sz, sz, szp, szp]),
        "ipb_rotatecrop_reset": (None, [vp]),
        "ipb_transform_transform_forward": (None, [vp, sz, sz, szp, szp]),
        "ipb_scaling_size": (None, [sz, sz, sz, sz, szp, szp]),
        "ipb_calculate_scale": (f, [sz, sz, sz, sz]),
        "ipb_spline_eval": (i, [vp, vp, vp, vp, sz]),
        "ipb_pack_8bit": (i, [vp, vp, vp, i]),
        "ipb_pack_16bit": (i, [vp, vp, vp, i]),
        "ipb_scale_down_srgb": (i, [vp, vp, sz, sz, sz, sz, vp, i]),
        "ipb_scale_down_srgb16": (i, [vp, vp, sz, sz, sz, sz, vp, i]),
        "ipb_lanczos_resize": (i, [vp, vp, sz, sz, i, vpp]),
        "ipb_ops_default": (None, [vp, vp]),
        "ipb_pipeline_create": (i, [vp, vp, vp, vpp]),
        "ipb_pipeline_destroy": (None, [vp]),
        "ipb_pipeline_ops": (vp, [vp]),
        "ipb_pipeline_settings": (vp, [vp]),
        "ipb_pipeline_set_source": (i, [vp, vp]),
        "ipb_pipeline_set_fused": (i, [vp, i]),
        "ipb_pipeline_output_size": (i, [vp, szp, szp]),
        "ipb_pipeline_run": (i, [vp, vpp]),
        "ipb_cache_create": (i, [vp, sz, vpp]),
        "ipb_cache_destroy": (None, [vp]),
        "ipb_cache_clear": (None, [vp]),
        "ipb_cache_bytes": (sz, [vp]),
        "ipb_cache_entries": (sz, [vp]),
        "ipb_pipeline_run_cached": (i, [vp, vp, vpp]),
        "ipb_pipeline_last_run_info": (None, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "ipb_pipeline_output_8bit": (i, [vp, vp, sz, i, szp, szp]),
        "ipb_pipeline_output_16bit": (i, [vp, vp, sz, i, szp, szp]),
        "ipb_pipeline_output_8bit_cached": (i, [vp, vp, vp, sz, i, szp, szp]),
        "ipb_pipeline_output_16bit_cached": (i, [vp, vp, vp, sz, i, szp, szp]),
        "ipb_pipeline_stripe_rows": (i, [vp, sz, sz, szp, szp]),
        "ipb_stripe_plan": (i, [vp, vp, sz, sz, sz, sz, szp, szp, szp, szp]),
        "ipb_pipeline_set_stripe_source": (i, [vp, vp, vp]),
        "ipb_pipeline_output_8bit_stripe": (i, [vp, vp, sz, i, szp, szp]),
        "ipb_pipeline_output_8bit_batch": (i, [vp, sz, sz, vp, sz, sz, szp, szp]),
        "ipb_pipeline_set_tma": (i, [vp, i]),
        "ipb_pipeline_set_band_mb": (i, [vp, i]),
        "ipb_pipeline_set_speculative": (i, [vp, i]),
        "ipb_comm_unique_id": (i, [C.c_char_p]),
        "ipb_comm_create": (i, [i, vp, C.c_char_p, i, i, C.POINTER(vp)]),
        "ipb_comm_destroy": (None, [vp]),
        "ipb_comm_rank": (i, [vp]),
        "ipb_comm_size": (i, [vp]),
        "ipb_comm_nccl_version": (i, [C.POINTER(i)]),
        "ipb_comm_last_error": (C.c_char_p, [vp]),
        "ipb_halo_exchange": (i, [vp, C.POINTER(vp), sz, C.POINTER(Halo)]),
        "ipb_ctx_set_spec": (i, [vp, C.c_float, i]),
        "ipb_spec_bound": (i, [vp, C.c_float, C.POINTER(C.c_float)]),
        "ipb_scaled_division_check": (i, [sz, sz, sz, sz]),
        "ipb_spec_tables": (i, [vp, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp]),
        "ipb_ctx_spec_stats": (i, [vp, C.POINTER(C.c_ulonglong), i]),
        "ipb_pipeline_spec_probe": (i, [vp, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_float)]),
        "ipb_selftest_gamma8": (i, [vp, C.POINTER(C.c_ulonglong)]),
        "ipb_gamma_pack_8bit": (i, [vp, vp, sz, vp]),
        "ipb_synth_cfa_u16": (i, [vp, C.c_uint64, sz, sz, sz, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._ipb_signatures = sigs
    _lib = L
    return L


def check(ctx_handle, rc):
    if rc != IPB_OK:
        msg = lib().ipb_last_error(ctx_handle)
        raise IpbError(rc, msg.decode() if msg else "")
