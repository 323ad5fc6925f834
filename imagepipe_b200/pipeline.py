"""Host-side mirror of the reference's Pipeline / ImageOp surface for the raw->sRGB hot path.

Names, argument meaning and error behaviour follow pedrocr/imagepipe (paths relative to the reference):
  ImageOp::run(&self, &PipelineGlobals, Arc<OpBuffer>) -> Arc<OpBuffer>     src/pipeline.rs:82-108
  PipelineOps{gofloat,demosaic,rotatecrop,tolab,basecurve,fromlab,gamma,transform}  src/pipeline.rs:154-164
  Pipeline::{new_from_source, run, output_8bit, output_16bit, default_ops}  src/pipeline.rs:257-470
Every op is a thin call into libipb200.so (CUDA, sm_100a); buffers stay on the GPU between ops.
There is no Python/NumPy implementation of any op here and no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import IpbError, lib  # noqa: F401  (re-exported)

Rotation = type("Rotation", (), {"Normal": 0, "Rotate90": 1, "Rotate180": 2, "Rotate270": 3})


# --------------------------------------------------------------------------------------------- context

class Context:
    """ipb_ctx: one device + one CUDA stream + the uploaded lookup tables."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        rc = lib().ipb_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        _capi.check(None, rc)
        self.handle = h
        self.device = device

    def synchronize(self):
        _capi.check(self.handle, lib().ipb_ctx_synchronize(self.handle))

    def set_stream(self, stream):
        _capi.check(self.handle, lib().ipb_ctx_set_stream(self.handle, C.c_void_p(stream) if stream else None))

    def set_spec(self, delta=0.0, threads=512):
        """Test / tuning hook of the speculative 8-bit kernel: forced bound (0 = certified) and CTA size."""
        _capi.check(self.handle, lib().ipb_ctx_set_spec(self.handle, float(delta), int(threads)))

    def spec_stats(self, reset=False):
        """dict(fixups, bad_window, delta, mufu_err) of the speculative kernel since the last reset."""
        import struct
        out = (C.c_ulonglong * 4)()
        _capi.check(self.handle, lib().ipb_ctx_spec_stats(self.handle, out, int(reset)))
        f = lambda v: struct.unpack("<f", struct.pack("<I", v & 0xffffffff))[0]
        return {"fixups": int(out[0]), "bad_window": int(out[1]), "delta": f(out[2]), "mufu_err": f(out[3])}

    @property
    def launch_count(self):
        return int(lib().ipb_ctx_launch_count(self.handle))

    def close(self):
        if self.handle:
            lib().ipb_ctx_destroy(self.handle)
            self.handle = None


_default_ctx = {}


def default_context(device=0):
    ctx = _default_ctx.get(device)
    if ctx is None or not ctx.handle:
        ctx = _default_ctx[device] = Context(device)
    return ctx


class DeviceArray:
    """Plain device memory (sources / destinations that live on the GPU)."""

    def __init__(self, nbytes, ctx=None):
        self.ctx = ctx or default_context()
        p = C.c_void_p()
        _capi.check(self.ctx.handle, lib().ipb_device_alloc(self.ctx.handle, nbytes, C.byref(p)))
        self.ptr = p.value
        self.nbytes = nbytes

    @classmethod
    def from_numpy(cls, arr, ctx=None):
        a = np.ascontiguousarray(arr)
        d = cls(a.nbytes, ctx)
        d.shape, d.dtype = a.shape, a.dtype
        _capi.check(d.ctx.handle, lib().ipb_device_upload(d.ctx.handle, d.ptr, a.ctypes.data, a.nbytes))
        return d

    def to_numpy(self, dtype=None, shape=None):
        dtype = np.dtype(dtype or self.dtype)
        out = np.empty(self.nbytes // dtype.itemsize, dtype)
        _capi.check(self.ctx.handle, lib().ipb_device_download(self.ctx.handle, out.ctypes.data, self.ptr, self.nbytes))
        shape = shape or getattr(self, "shape", None)
        return out.reshape(shape) if shape else out

    def free(self):
        if self.ptr and self.ctx.handle:
            lib().ipb_device_free(self.ctx.handle, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------- OpBuffer

class OpBuffer:
    """Device-resident OpBuffer (src/buffer.rs:4-11): interleaved row-major f32, ref-counted like Arc."""

    def __init__(self, handle, ctx):
        self.handle = handle
        self.ctx = ctx

    @classmethod
    def new(cls, width, height, colors, monochrome=False, ctx=None):
        ctx = ctx or default_context()
        h = C.c_void_p()
        _capi.check(ctx.handle, lib().ipb_buffer_new(ctx.handle, width, height, colors, int(monochrome), C.byref(h)))
        return cls(h, ctx)

    @classmethod
    def from_numpy(cls, arr, monochrome=False, ctx=None):
        ctx = ctx or default_context()
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if a.ndim == 2:
            a = a[:, :, None]
        hgt, wid, col = a.shape
        h = C.c_void_p()
        _capi.check(ctx.handle, lib().ipb_buffer_upload(ctx.handle, wid, hgt, col, int(monochrome), a.ctypes.data, C.byref(h)))
        return cls(h, ctx)

    @classmethod
    def from_rgb_str_vec(cls, rows, ctx=None):
        """buffer.rs:82-113: human-readable 3-channel buffers ('R','G','B','O',' ')."""
        table = {"R": (1, 0, 0), "G": (0, 1, 0), "B": (0, 0, 1), "O": (1, 1, 1), " ": (0, 0, 0)}
        a = np.array([[table[c] for c in row] for row in rows], np.float32)
        return cls.from_numpy(a, ctx=ctx)

    width = property(lambda s: int(lib().ipb_buffer_width(s.handle)))
    height = property(lambda s: int(lib().ipb_buffer_height(s.handle)))
    colors = property(lambda s: int(lib().ipb_buffer_colors(s.handle)))
    monochrome = property(lambda s: bool(lib().ipb_buffer_monochrome(s.handle)))
    device_ptr = property(lambda s: lib().ipb_buffer_device_ptr(s.handle))

    def to_numpy(self):
        out = np.empty((self.height, self.width, self.colors), np.float32)
        _capi.check(self.ctx.handle, lib().ipb_buffer_download(self.ctx.handle, self.handle, out.ctypes.data))
        return out

    @property
    def data(self):
        return self.to_numpy().reshape(-1)

    def same_arc(self, other):
        """True when both wrappers hold the same underlying buffer (the reference returns the same Arc)."""
        return self.handle.value == other.handle.value

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                lib().ipb_buffer_release(self.handle)
        except Exception:
            pass
        self.handle = None


def _run_op(ctx, fn, *args):
    out = C.c_void_p()
    _capi.check(ctx.handle, fn(ctx.handle, *args, C.byref(out)))
    return OpBuffer(out, ctx)


# --------------------------------------------------------------------------------------------- sources

class ImageSource:
    """ImageSource::{Raw, Other} (src/pipeline.rs:46-50) with literal metadata instead of a decoded file."""

    def __init__(self, kind, width, height, cpp, data, keep=None):
        self.kind, self.width, self.height, self.cpp = kind, width, height, cpp
        self._keep = keep if keep is not None else data
        on_device = isinstance(data, DeviceArray) or isinstance(data, int)
        ptr = data.ptr if isinstance(data, DeviceArray) else (data if isinstance(data, int) else data.ctypes.data)
        self.c = _capi.Source(kind, width, height, cpp, ptr, int(on_device))

    @classmethod
    def Raw(cls, data, width=None, height=None, cpp=1, dtype=None):
        """rawloader::RawImage{width,height,cpp,data}: u16 (Integer) or f32 (Float) samples."""
        if isinstance(data, np.ndarray):
            a = np.ascontiguousarray(data)
            if a.dtype not in (np.uint16, np.float32):
                raise TypeError("raw data must be uint16 or float32")
            height, width = a.shape[:2]
            cpp = a.shape[2] if a.ndim == 3 else 1
            kind = _capi.SRC_RAW_U16 if a.dtype == np.uint16 else _capi.SRC_RAW_F32
            return cls(kind, width, height, cpp, a)
        kind = _capi.SRC_RAW_F32 if np.dtype(dtype or np.uint16) == np.float32 else _capi.SRC_RAW_U16
        return cls(kind, width, height, cpp, data)

    @classmethod
    def Other(cls, data, width=None, height=None, dtype=None):
        """image::DynamicImage::ImageRgb8 / ImageRgb16 raster."""
        if isinstance(data, np.ndarray):
            a = np.ascontiguousarray(data)
            if a.dtype not in (np.uint8, np.uint16) or a.ndim != 3 or a.shape[2] != 3:
                raise TypeError("raster must be (h, w, 3) uint8 or uint16")
            kind = _capi.SRC_RGB8 if a.dtype == np.uint8 else _capi.SRC_RGB16
            return cls(kind, a.shape[1], a.shape[0], 3, a)
        kind = _capi.SRC_RGB16 if np.dtype(dtype or np.uint8) == np.uint16 else _capi.SRC_RGB8
        return cls(kind, width, height, 3, data)

    @property
    def is_raw(self):
        return self.kind in (_capi.SRC_RAW_U16, _capi.SRC_RAW_F32)


# --------------------------------------------------------------------------------------------- ops

class PipelineSettings(_capi.Settings):
    """src/pipeline.rs:110-131"""

    @classmethod
    def default(cls):
        return cls(0, 0, 0, 0, 0, 1)


class PipelineGlobals:
    """src/pipeline.rs:139-152"""

    def __init__(self, image, settings=None, ctx=None):
        self.image = image
        self.settings = settings if settings is not None else PipelineSettings.default()
        self.ctx = ctx or default_context()

    @classmethod
    def mock(cls, width, height, ctx=None):
        return cls(ImageSource.Other(np.zeros((height, width, 3), np.uint8)), ctx=ctx)


class _SizeMixin:
    def transform_forward(self, width, height):
        return (width, height)

    def transform_reverse(self, width, height):
        return (width, height)

    def reset(self):
        pass


def _wh(fn, op, w, h):
    ow, oh = C.c_size_t(), C.c_size_t()
    fn(C.byref(op), w, h, C.byref(ow), C.byref(oh))
    return ow.value, oh.value


class OpGoFloat(_capi.GoFloat, _SizeMixin):
    """src/ops/gofloat.rs"""
    name = "gofloat"

    def run(self, pipeline, buf=None):
        return _run_op(pipeline.ctx, lib().ipb_gofloat_run, C.byref(self), C.byref(pipeline.image.c))

    def transform_forward(self, width, height):
        return _wh(lib().ipb_gofloat_transform_forward, self, width, height)


class OpDemosaic(_capi.Demosaic, _SizeMixin):
    """src/ops/demosaic.rs"""
    name = "demosaic"

    def run(self, pipeline, buf):
        return _run_op(pipeline.ctx, lib().ipb_demosaic_run, C.byref(self), C.byref(pipeline.settings), buf.handle)


class OpRotateCrop(_capi.RotateCrop, _SizeMixin):
    """src/ops/rotatecrop.rs"""
    name = "rotatecrop"

    @classmethod
    def empty(cls):
        op = cls()
        op.input_ratio = 1.0
        return op

    def run(self, pipeline, buf):
        return _run_op(pipeline.ctx, lib().ipb_rotatecrop_run, C.byref(self), buf.handle)

    def transform_forward(self, width, height):
        return _wh(lib().ipb_rotatecrop_transform_forward, self, width, height)

    def transform_reverse(self, width, height):
        return _wh(lib().ipb_rotatecrop_transform_reverse, self, width, height)

    def reset(self):
        lib().ipb_rotatecrop_reset(C.byref(self))


class OpToLab(_capi.ToLab, _SizeMixin):
    """src/ops/colorspaces.rs:5-113"""
    name = "to_lab"

    def run(self, pipeline, buf):
        return _run_op(pipeline.ctx, lib().ipb_tolab_run, C.byref(self), buf.handle)


class OpBaseCurve(_capi.BaseCurve, _SizeMixin):
    """src/ops/curves.rs"""
    name = "basecurve"

    def set_points(self, pts):
        if len(pts) > _capi.MAX_CURVE_POINTS:
            raise ValueError("too many curve points")
        self.npoints = len(pts)
        for i, (x, y) in enumerate(pts):
            self.points[i][0] = x
            self.points[i][1] = y

    def get_points(self):
        return [(self.points[i][0], self.points[i][1]) for i in range(self.npoints)]

    def run(self, pipeline, buf):
        return _run_op(pipeline.ctx, lib().ipb_basecurve_run, C.byref(self), buf.handle)


class SplineFunc:
    """SplineFunc::new(points).interpolate(v) (src/ops/curves.rs:59-157), evaluated by the device kernel."""

    def __init__(self, points, ctx=None):
        self.ctx = ctx or default_context()
        self.op = OpBaseCurve()
        self.op.set_points(points)

    def interpolate(self, val):
        vals = np.atleast_1d(np.asarray(val, np.float32))
        out = np.empty_like(vals)
        _capi.check(self.ctx.handle, lib().ipb_spline_eval(self.ctx.handle, C.byref(self.op), vals.ctypes.data,
                                                            out.ctypes.data, vals.size))
        return float(out[0]) if np.isscalar(val) else out


class OpFromLab(_SizeMixin):
    """src/ops/colorspaces.rs:115-138"""
    name = "from_lab"

    def run(self, pipeline, buf):
        return _run_op(pipeline.ctx, lib().ipb_fromlab_run, buf.handle)


class OpGamma(_SizeMixin):
    """src/ops/gamma.rs"""
    name = "gamma"

    def run(self, pipeline, buf):
        return _run_op(pipeline.ctx, lib().ipb_gamma_run, C.byref(pipeline.settings), buf.handle)


class OpTransform(_capi.Transform, _SizeMixin):
    """src/ops/transform.rs"""
    name = "transform"

    def run(self, pipeline, buf):
        return _run_op(pipeline.ctx, lib().ipb_transform_run, C.byref(self), buf.handle)

    def transform_forward(self, width, height):
        return _wh(lib().ipb_transform_transform_forward, self, width, height)

    transform_reverse = transform_forward


# rawloader Orientation::to_flips -> (transpose, flip_x, flip_y); pinned by the golden bitmaps of transform.rs:167-278
ORIENTATION_FLIPS = {
    "Normal": (False, False, False), "Unknown": (False, False, False), "VerticalFlip": (False, False, True),
    "HorizontalFlip": (False, True, False), "Rotate180": (False, True, True), "Transpose": (True, False, False),
    "Rotate90": (True, False, True), "Rotate270": (True, True, False), "Transverse": (True, True, True),
}


def rotate_buffer(buf, orientation):
    """rotate_buffer(buf, &orientation) of transform.rs:87-144 for a named rawloader Orientation, expressed through
    the OpTransform fields whose run() derives exactly these flips (transform.rs:56-66).  Normal / Unknown pass
    the same buffer through (the reference's rotate_buffer clones it; the values are identical)."""
    t, fx, fy = ORIENTATION_FLIPS[orientation]
    rot, fh, fv = (Rotation.Rotate90, fx, not fy) if t else (Rotation.Normal, fx, fy)
    op = OpTransform(rot, int(fh), int(fv))
    g = PipelineGlobals.mock(16, 16, ctx=buf.ctx)
    return op.run(g, buf)


class PipelineOps(C.Structure):
    """src/pipeline.rs:154-164 — same memory layout as ipb_ops; fromlab/gamma have no fields."""
    _fields_ = [("gofloat", OpGoFloat), ("demosaic", OpDemosaic), ("rotatecrop", OpRotateCrop), ("tolab", OpToLab),
                ("basecurve", OpBaseCurve), ("transform", OpTransform)]
    fromlab = OpFromLab()
    gamma = OpGamma()

    @classmethod
    def new(cls, img):
        ops = cls()
        lib().ipb_ops_default(C.byref(ops), C.byref(img.c))
        return ops

    def all_ops(self):
        return [self.gofloat, self.demosaic, self.rotatecrop, self.tolab, self.basecurve, self.fromlab, self.gamma,
                self.transform]


class SRGBImage:
    """src/pipeline.rs:26-41 (SRGBImage / SRGBImage16)"""

    def __init__(self, width, height, data):
        self.width, self.height, self.data = width, height, data

    def to_numpy(self):
        return self.data.reshape(self.height, self.width, 3)


SRGBImage16 = SRGBImage


# --------------------------------------------------------------------------------------------- Pipeline

class PipelineCache:
    """PipelineCache = MultiCache<BufHash, OpBuffer> (src/pipeline.rs:43), holding device buffers."""

    def __init__(self, size, ctx=None):
        self.ctx = ctx or default_context()
        self.handle = C.c_void_p()
        _capi.check(self.ctx.handle, lib().ipb_cache_create(self.ctx.handle, int(size), C.byref(self.handle)))

    bytes = property(lambda s: int(lib().ipb_cache_bytes(s.handle)))
    entries = property(lambda s: int(lib().ipb_cache_entries(s.handle)))

    def clear(self):
        lib().ipb_cache_clear(self.handle)

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                lib().ipb_cache_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


class Pipeline:
    """src/pipeline.rs:245-470.  `globals.settings` and `ops` are live views of the native pipeline object."""

    def __init__(self, img, ops=None, ctx=None):
        self.ctx = ctx or default_context()
        self.handle = C.c_void_p()
        _capi.check(self.ctx.handle, lib().ipb_pipeline_create(self.ctx.handle, C.byref(img.c),
                                                              C.byref(ops) if ops is not None else None,
                                                              C.byref(self.handle)))
        self.ops = C.cast(lib().ipb_pipeline_ops(self.handle), C.POINTER(PipelineOps)).contents
        settings = C.cast(lib().ipb_pipeline_settings(self.handle), C.POINTER(PipelineSettings)).contents
        self.globals = PipelineGlobals(img, settings, self.ctx)

    @classmethod
    def new_from_source(cls, img, ctx=None):
        return cls(img, ctx=ctx)

    def set_source(self, img):
        """Swap the pixel data (same metadata): the next frame of a batch."""
        _capi.check(self.ctx.handle, lib().ipb_pipeline_set_source(self.handle, C.byref(img.c)))
        self.globals.image = img

    def set_fused(self, fused):
        """True (default): run may use the fused raw->sRGB kernel; False: one kernel + one OpBuffer per op."""
        _capi.check(self.ctx.handle, lib().ipb_pipeline_set_fused(self.handle, int(fused)))

    def set_speculative(self, on):
        """True (default): 8-bit output of RGB Bayer frames through the speculative kernel (identical bytes)."""
        _capi.check(self.ctx.handle, lib().ipb_pipeline_set_speculative(self.handle, int(on)))

    def spec_probe(self):
        """(max |cheap - exact|, mean, certified delta) over the linear channel values of this frame's interior pixels."""
        mx, mean, delta = C.c_float(), C.c_double(), C.c_float()
        _capi.check(self.ctx.handle, lib().ipb_pipeline_spec_probe(self.handle, C.byref(mx), C.byref(mean), C.byref(delta)))
        return mx.value, mean.value, delta.value

    def set_band_mb(self, megabytes):
        """Host source/destination: band size of the overlapped H2D / kernel / D2H schedule (0 = whole-frame copies)."""
        _capi.check(self.ctx.handle, lib().ipb_pipeline_set_band_mb(self.handle, int(megabytes)))

    def default_ops(self):
        return bytes(self.ops) == bytes(PipelineOps.new(self.globals.image))

    def output_size(self):
        w, h = C.c_size_t(), C.c_size_t()
        _capi.check(self.ctx.handle, lib().ipb_pipeline_output_size(self.handle, C.byref(w), C.byref(h)))
        return w.value, h.value

    @staticmethod
    def new_cache(size, ctx=None):
        """Pipeline::new_cache(size) (pipeline.rs:257-260): a device-resident LRU of op outputs, `size` bytes."""
        return PipelineCache(size, ctx)

    def run(self, cache=None):
        """Pipeline::run(cache) (pipeline.rs:311-375).  With a cache the ops run one by one from the first op whose
        parameters changed since a cached run; without, the fused kernel is used when the chain allows it."""
        out = C.c_void_p()
        if cache is not None:
            _capi.check(self.ctx.handle, lib().ipb_pipeline_run_cached(self.handle, cache.handle, C.byref(out)))
        else:
            _capi.check(self.ctx.handle, lib().ipb_pipeline_run(self.handle, C.byref(out)))
        return OpBuffer(out, self.ctx)

    def last_run_info(self):
        """(index of the first op the last cached run executed — 8: none —, number of ops executed)"""
        a, b = C.c_int(), C.c_int()
        lib().ipb_pipeline_last_run_info(self.handle, C.byref(a), C.byref(b))
        return a.value, b.value

    def _output_cached(self, fn, cache, dtype):
        # pipeline.rs:377-422 / :424-469 with Some(&cache): fast path first, else settings.linear, run(cache), pack loop
        w, h = self.output_size()
        host = np.empty(max(w * h, self.globals.image.width * self.globals.image.height) * 3, dtype)
        ow, oh = C.c_size_t(), C.c_size_t()
        _capi.check(self.ctx.handle, fn(self.handle, cache.handle, host.ctypes.data, host.size, 0, C.byref(ow), C.byref(oh)))
        return SRGBImage(ow.value, oh.value, host[: ow.value * oh.value * 3])

    def _output(self, fn, dtype, dst):
        w, h = self.output_size()
        cap = w * h * 3
        if dst is None:
            host = np.empty(cap, dtype)
            ptr, on_device = host.ctypes.data, 0
        elif isinstance(dst, np.ndarray):
            host, ptr, on_device = dst.reshape(-1), dst.ctypes.data, 0
            cap = host.size
        else:
            host, ptr, on_device = dst, dst.ptr, 1
            cap = dst.nbytes // np.dtype(dtype).itemsize
        ow, oh = C.c_size_t(), C.c_size_t()
        _capi.check(self.ctx.handle, fn(self.handle, ptr, cap, on_device, C.byref(ow), C.byref(oh)))
        if not on_device:
            host = host[: ow.value * oh.value * 3]
        return SRGBImage(ow.value, oh.value, host)

    def output_8bit(self, cache=None, dst=None):
        if cache is not None:
            return self._output_cached(lib().ipb_pipeline_output_8bit_cached, cache, np.uint8)
        return self._output(lib().ipb_pipeline_output_8bit, np.uint8, dst)

    def output_16bit(self, cache=None, dst=None):
        if cache is not None:
            return self._output_cached(lib().ipb_pipeline_output_16bit_cached, cache, np.uint16)
        return self._output(lib().ipb_pipeline_output_16bit, np.uint16, dst)

    # ---- row stripes (multi-GPU sharding of one large frame; no reference equivalent)
    def stripe_rows(self, out_row0, out_row1):
        a, b = C.c_size_t(), C.c_size_t()
        _capi.check(self.ctx.handle, lib().ipb_pipeline_stripe_rows(self.handle, out_row0, out_row1, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_stripe_source(self, rows_src, src_row0, out_row0, out_row1):
        st = _capi.Stripe(self.globals.image.height, src_row0, out_row0, out_row1)
        _capi.check(self.ctx.handle, lib().ipb_pipeline_set_stripe_source(self.handle, C.byref(rows_src.c), C.byref(st)))
        self._stripe_keep = rows_src

    def output_8bit_stripe(self, dst=None, rows=None, width=None):
        if dst is None:
            host = np.empty(rows * width * 3, np.uint8)
            ptr, on_device, cap = host.ctypes.data, 0, host.size
        elif isinstance(dst, np.ndarray):
            host, ptr, on_device, cap = dst.reshape(-1), dst.ctypes.data, 0, dst.size
        else:
            host, ptr, on_device, cap = dst, dst.ptr, 1, dst.nbytes
        ow, orows = C.c_size_t(), C.c_size_t()
        _capi.check(self.ctx.handle, lib().ipb_pipeline_output_8bit_stripe(self.handle, ptr, cap, on_device, C.byref(ow),
                                                                        C.byref(orows)))
        if not on_device:
            host = host[: ow.value * orows.value * 3]
        return SRGBImage(ow.value, orows.value, host)

    def output_8bit_batch(self, nframes, src_stride_rows, dst, dst_stride_bytes):
        """ipb_pipeline_output_8bit_batch: nframes device-resident frames of this pipeline's geometry, frame k's source
        rows src_stride_rows * k rows after the pipeline's source (image or stripe rows), its result dst_stride_bytes * k
        bytes into the DeviceArray dst.  Returns (width, rows) of one result."""
        ow, orows = C.c_size_t(), C.c_size_t()
        _capi.check(self.ctx.handle, lib().ipb_pipeline_output_8bit_batch(self.handle, nframes, src_stride_rows, dst.ptr,
                                                                       dst_stride_bytes, dst.nbytes, C.byref(ow), C.byref(orows)))
        return ow.value, orows.value

    def close(self):
        if self.handle and self.ctx.handle:
            lib().ipb_pipeline_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scale_down_srgb(img, nwidth, nheight, ctx=None):
    """scaling.rs:162-182 for (h, w, 3) uint8 / uint16 rasters."""
    ctx = ctx or default_context()
    a = np.ascontiguousarray(img)
    out = np.empty((nheight, nwidth, 3), a.dtype)
    fn = lib().ipb_scale_down_srgb if a.dtype == np.uint8 else lib().ipb_scale_down_srgb16
    _capi.check(ctx.handle, fn(ctx.handle, a.ctypes.data, a.shape[1], a.shape[0], nwidth, nheight, out.ctypes.data, 0))
    return out


def lanczos_resize(buf, nwidth, nheight, a=3):
    """EXTENSION: Lanczos-a separable resample of a device OpBuffer (the reference has none: scaling.rs:101-103)."""
    return _run_op(buf.ctx, lib().ipb_lanczos_resize, buf.handle, int(nwidth), int(nheight), int(a))


def scaling_size(width, height, maxwidth, maxheight):
    ow, oh = C.c_size_t(), C.c_size_t()
    lib().ipb_scaling_size(width, height, maxwidth, maxheight, C.byref(ow), C.byref(oh))
    return ow.value, oh.value


def calculate_scale(width, height, maxwidth, maxheight):
    return float(lib().ipb_calculate_scale(width, height, maxwidth, maxheight))


def synth_cfa_u16(seed, width, row0, rows, ctx=None):
    """SURVEY.md §8d synthetic frames: v(i) = splitmix64(seed ^ i) mod 16384, generated on the device."""
    ctx = ctx or default_context()
    d = DeviceArray(width * rows * 2, ctx)
    d.shape, d.dtype = (rows, width), np.dtype(np.uint16)
    _capi.check(ctx.handle, lib().ipb_synth_cfa_u16(ctx.handle, seed, width, row0, rows, d.ptr))
    return d
