// gpu.rs — the reference-side shim over libipb200.so: what a maintainer of pedrocr/imagepipe adds as `src/gpu.rs`
// (+ `mod gpu;` in src/lib.rs and the generated `ffi.rs` next to it) to run the OpBuffer hot path on a B200.
//
// NOT compiled in this repository: the image has no rustc / cargo, and the reference's path dependencies (rawloader,
// multicache) are absent.  It is written against the reference's own types (src/pipeline.rs, src/ops/*.rs,
// src/buffer.rs @ 65ca96ce) and the generated ffi.rs; tests/test_shim_rs.py checks that every C symbol, struct field
// and field order it relies on exists in include/ipb200.h.
//
// Shape of the integration:
//   * `GpuContext`            one per device + stream (ipb_ctx); cheap to share by reference.
//   * `GpuBuffer`             a device-resident OpBuffer with Arc semantics (ipb_buffer_retain / release) — the
//                             `Arc<OpBuffer>` of the reference while the data stays on the GPU.
//   * `Gpu<Op>`               newtype over each of the reference's eight ops; `impl ImageOp for Gpu<Op>` keeps the
//                             trait's names, argument meaning and error behaviour (ops are infallible by signature:
//                             a CUDA failure panics with ipb_last_error, like the reference's assert_eq! panics).
//                             `run` takes and returns host `Arc<OpBuffer>` as the trait demands (upload, kernel,
//                             download); `run_device` is the zero-copy form the pipeline hand-off uses.
//   * `GpuPipeline`           the hand-off of `Pipeline::run` / `output_8bit` / `output_16bit`
//                             (src/pipeline.rs:311-469): the whole op chain in one C call — one fused kernel when the
//                             chain allows it — including the size negotiation, with the cache variant.
use std::ffi::CStr;
use std::os::raw::c_int;
use std::ptr;
use std::sync::Arc;

use crate::buffer::OpBuffer;
use crate::gpu::ffi::*;
use crate::ops::colorspaces::{OpFromLab, OpToLab};
use crate::ops::curves::OpBaseCurve;
use crate::ops::demosaic::OpDemosaic;
use crate::ops::gamma::OpGamma;
use crate::ops::gofloat::OpGoFloat;
use crate::ops::rotatecrop::OpRotateCrop;
use crate::ops::transform::{OpTransform, Rotation};
use crate::pipeline::{ImageOp, ImageSource, Pipeline, PipelineGlobals, PipelineOps, PipelineSettings, SRGBImage, SRGBImage16};

pub mod ffi;

// ------------------------------------------------------------------------------------------------ context

pub struct GpuContext {
    raw: *mut ipb_ctx,
}
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {} // calls on one context are stream-ordered; distinct contexts are independent

impl GpuContext {
    /// `stream`: a cudaStream_t to run on, or null for a private stream (ipb_ctx_create).
    pub fn new(device: i32, stream: *mut std::os::raw::c_void) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { ipb_ctx_create(device as c_int, stream, &mut raw) };
        if rc != IPB_OK {
            return Err(last_error(ptr::null()));
        }
        Ok(Self { raw })
    }
    pub fn synchronize(&self) {
        self.check(unsafe { ipb_ctx_synchronize(self.raw) });
    }
    fn check(&self, rc: c_int) {
        // ops are infallible by signature (src/pipeline.rs:84): a failing launch panics like the reference's asserts
        if rc != IPB_OK {
            panic!("imagepipe-b200: {}", last_error(self.raw));
        }
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { ipb_ctx_destroy(self.raw) }
    }
}
fn last_error(ctx: *const ipb_ctx) -> String {
    unsafe { CStr::from_ptr(ipb_last_error(ctx)).to_string_lossy().into_owned() }
}

// ------------------------------------------------------------------------------------------------ device OpBuffer

/// `Arc<OpBuffer>` on the device: interleaved row-major f32, width * height * colors elements (src/buffer.rs:4-11).
pub struct GpuBuffer {
    raw: *mut ipb_buffer,
}
unsafe impl Send for GpuBuffer {}
unsafe impl Sync for GpuBuffer {}

impl GpuBuffer {
    pub fn upload(ctx: &GpuContext, buf: &OpBuffer) -> Self {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe {
            ipb_buffer_upload(ctx.raw, buf.width, buf.height, buf.colors, buf.monochrome as c_int, buf.data.as_ptr(), &mut raw)
        });
        Self { raw }
    }
    pub fn download(&self, ctx: &GpuContext) -> OpBuffer {
        let (width, height, colors) = unsafe { (ipb_buffer_width(self.raw), ipb_buffer_height(self.raw), ipb_buffer_colors(self.raw)) };
        let mut out = OpBuffer::new(width, height, colors, unsafe { ipb_buffer_monochrome(self.raw) } != 0);
        ctx.check(unsafe { ipb_buffer_download(ctx.raw, self.raw, out.data.as_mut_ptr()) });
        out
    }
    /// true when both handles name the same device buffer — what `Arc::ptr_eq` is for pass-through ops
    /// (demosaic.rs:43, rotatecrop.rs:40, curves.rs:35, gamma.rs:18, transform.rs:69)
    pub fn same_arc(&self, other: &GpuBuffer) -> bool {
        self.raw == other.raw
    }
}
impl Clone for GpuBuffer {
    fn clone(&self) -> Self {
        unsafe { ipb_buffer_retain(self.raw) };
        Self { raw: self.raw }
    }
}
impl Drop for GpuBuffer {
    fn drop(&mut self) {
        unsafe { ipb_buffer_release(self.raw) }
    }
}

// ------------------------------------------------------------------------------------------------ parameter twins

fn pod_gofloat(op: &OpGoFloat) -> ipb_gofloat {
    ipb_gofloat {
        crop_top: op.crop_top, crop_right: op.crop_right, crop_bottom: op.crop_bottom, crop_left: op.crop_left,
        is_cfa: op.is_cfa as c_int, blacklevels: op.blacklevels, whitelevels: op.whitelevels,
    }
}
fn pod_demosaic(op: &OpDemosaic) -> ipb_demosaic {
    let mut cfa = [0 as std::os::raw::c_char; 148];
    for (dst, src) in cfa.iter_mut().zip(op.cfa.bytes().take(147)) {
        *dst = src as std::os::raw::c_char;
    }
    ipb_demosaic { cfa }
}
fn pod_rotatecrop(op: &OpRotateCrop) -> ipb_rotatecrop {
    // input_ratio / output_size are private in the reference (rotatecrop.rs:16-17): the shim lives in the same crate,
    // or the two fields gain `pub(crate)`
    let (has, (w, h)) = match op.output_size { Some(s) => (1, s), None => (0, (0, 0)) };
    ipb_rotatecrop {
        crop_top: op.crop_top, crop_right: op.crop_right, crop_bottom: op.crop_bottom, crop_left: op.crop_left,
        rotation: op.rotation, input_ratio: op.input_ratio, has_output_size: has, output_width: w, output_height: h,
    }
}
fn unpod_rotatecrop(op: &mut OpRotateCrop, pod: &ipb_rotatecrop) {
    op.input_ratio = pod.input_ratio;
    op.output_size = if pod.has_output_size != 0 { Some((pod.output_width, pod.output_height)) } else { None };
}
fn pod_tolab(op: &OpToLab) -> ipb_tolab {
    ipb_tolab { cam_to_xyz: op.cam_to_xyz, cam_to_xyz_normalized: op.cam_to_xyz_normalized, xyz_to_cam: op.xyz_to_cam, wb_coeffs: op.wb_coeffs }
}
fn pod_basecurve(op: &OpBaseCurve) -> ipb_basecurve {
    assert!(op.points.len() <= IPB_MAX_CURVE_POINTS, "imagepipe-b200 carries at most {} curve points", IPB_MAX_CURVE_POINTS);
    let mut points = [[0.0f32; 2]; 32];
    for (dst, (x, y)) in points.iter_mut().zip(op.points.iter()) {
        *dst = [*x, *y];
    }
    ipb_basecurve { exposure: op.exposure, npoints: op.points.len(), points }
}
fn pod_transform(op: &OpTransform) -> ipb_transform {
    let rotation = match op.rotation {
        Rotation::Normal => IPB_ROT_NORMAL, Rotation::Rotate90 => IPB_ROT_90,
        Rotation::Rotate180 => IPB_ROT_180, Rotation::Rotate270 => IPB_ROT_270,
    };
    ipb_transform { rotation, fliph: op.fliph as c_int, flipv: op.flipv as c_int }
}
fn pod_settings(s: &PipelineSettings) -> ipb_settings {
    ipb_settings {
        maxwidth: s.maxwidth, maxheight: s.maxheight, demosaic_width: s.demosaic_width, demosaic_height: s.demosaic_height,
        linear: s.linear as c_int, use_fastpath: s.use_fastpath as c_int,
    }
}
fn pod_ops(ops: &PipelineOps) -> ipb_ops {
    ipb_ops {
        gofloat: pod_gofloat(&ops.gofloat), demosaic: pod_demosaic(&ops.demosaic), rotatecrop: pod_rotatecrop(&ops.rotatecrop),
        tolab: pod_tolab(&ops.tolab), basecurve: pod_basecurve(&ops.basecurve), transform: pod_transform(&ops.transform),
    }
}
/// ImageSource -> ipb_source (src/pipeline.rs:46-50).  The pixel data is NOT copied: `keep` owns what `data` points at
/// for the non-raw case (the `image` crate's to_rgb8 / to_rgb16 raster), the RawImage owns it otherwise.
pub struct SourceView {
    pub pod: ipb_source,
    keep: Option<Vec<u8>>,
    keep16: Option<Vec<u16>>,
}
pub fn source_view(img: &ImageSource) -> SourceView {
    use rawloader::RawImageData;
    match img {
        ImageSource::Raw(raw) => {
            let (kind, data) = match &raw.data {
                RawImageData::Integer(v) => (IPB_SRC_RAW_U16, v.as_ptr() as *const std::os::raw::c_void),
                RawImageData::Float(v) => (IPB_SRC_RAW_F32, v.as_ptr() as *const std::os::raw::c_void),
            };
            SourceView { pod: ipb_source { kind, width: raw.width, height: raw.height, cpp: raw.cpp, data, on_device: 0 }, keep: None, keep16: None }
        }
        ImageSource::Other(img) => {
            use image::DynamicImage;
            match img {
                DynamicImage::ImageRgb16(_) | DynamicImage::ImageRgba16(_) | DynamicImage::ImageLuma16(_) | DynamicImage::ImageLumaA16(_) => {
                    let raster = img.to_rgb16();
                    let (w, h) = (raster.width() as usize, raster.height() as usize);
                    let v = raster.into_raw();
                    let data = v.as_ptr() as *const std::os::raw::c_void;
                    SourceView { pod: ipb_source { kind: IPB_SRC_RGB16, width: w, height: h, cpp: 3, data, on_device: 0 }, keep: None, keep16: Some(v) }
                }
                _ => {
                    let raster = img.to_rgb8();
                    let (w, h) = (raster.width() as usize, raster.height() as usize);
                    let v = raster.into_raw();
                    let data = v.as_ptr() as *const std::os::raw::c_void;
                    SourceView { pod: ipb_source { kind: IPB_SRC_RGB8, width: w, height: h, cpp: 3, data, on_device: 0 }, keep: Some(v), keep16: None }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ the eight ops

/// A reference op bound to a GPU context.  `Gpu<OpX>` implements `ImageOp` with the reference's semantics; serde goes
/// through the wrapped op, so settings files and cache hashes do not change.
#[derive(Debug, Clone)]
pub struct Gpu<'c, Op> {
    pub op: Op,
    pub ctx: &'c GpuContext,
}
impl<'c, Op: serde::Serialize> serde::Serialize for Gpu<'c, Op> {
    fn serialize<S: serde::Serializer>(&self, s: S) -> Result<S::Ok, S::Error> {
        self.op.serialize(s)
    }
}
impl std::fmt::Debug for GpuContext {
    fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        write!(f, "GpuContext({:p})", self.raw)
    }
}

/// upload -> device op -> download: the host-buffer form the trait signature demands
fn through_device(ctx: &GpuContext, buf: Arc<OpBuffer>, f: impl FnOnce(&GpuBuffer) -> GpuBuffer) -> Arc<OpBuffer> {
    let dev_in = GpuBuffer::upload(ctx, &buf);
    let dev_out = f(&dev_in);
    if dev_out.same_arc(&dev_in) {
        return buf; // pass-through ops hand back the very same Arc, like the reference
    }
    Arc::new(dev_out.download(ctx))
}

macro_rules! device_call {
    ($ctx:expr, $call:ident ( $($arg:expr),* )) => {{
        let mut out = ptr::null_mut();
        $ctx.check(unsafe { $call($ctx.raw, $($arg,)* &mut out) });
        GpuBuffer { raw: out }
    }};
}

impl<'c> Gpu<'c, OpGoFloat> {
    /// gofloat reads the image, not `buf` (src/ops/gofloat.rs:50-62)
    pub fn run_device(&self, globals: &PipelineGlobals) -> GpuBuffer {
        let src = source_view(&globals.image);
        device_call!(self.ctx, ipb_gofloat_run(&pod_gofloat(&self.op), &src.pod))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpGoFloat> {
    fn name(&self) -> &str { "gofloat" }
    fn run(&self, globals: &PipelineGlobals, _buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        Arc::new(self.run_device(globals).download(self.ctx))
    }
    fn transform_forward(&mut self, width: usize, height: usize) -> (usize, usize) {
        let (mut w, mut h) = (0, 0);
        unsafe { ipb_gofloat_transform_forward(&pod_gofloat(&self.op), width, height, &mut w, &mut h) };
        (w, h)
    }
}

impl<'c> Gpu<'c, OpDemosaic> {
    pub fn run_device(&self, globals: &PipelineGlobals, buf: &GpuBuffer) -> GpuBuffer {
        device_call!(self.ctx, ipb_demosaic_run(&pod_demosaic(&self.op), &pod_settings(&globals.settings), buf.raw))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpDemosaic> {
    fn name(&self) -> &str { "demosaic" }
    fn run(&self, globals: &PipelineGlobals, buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        through_device(self.ctx, buf, |b| self.run_device(globals, b))
    }
}

impl<'c> Gpu<'c, OpRotateCrop> {
    pub fn run_device(&self, buf: &GpuBuffer) -> GpuBuffer {
        device_call!(self.ctx, ipb_rotatecrop_run(&pod_rotatecrop(&self.op), buf.raw))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpRotateCrop> {
    fn name(&self) -> &str { "rotatecrop" }
    fn run(&self, _globals: &PipelineGlobals, buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        through_device(self.ctx, buf, |b| self.run_device(b))
    }
    fn transform_forward(&mut self, width: usize, height: usize) -> (usize, usize) {
        let mut pod = pod_rotatecrop(&self.op);
        let (mut w, mut h) = (0, 0);
        unsafe { ipb_rotatecrop_transform_forward(&mut pod, width, height, &mut w, &mut h) };
        unpod_rotatecrop(&mut self.op, &pod); // the op remembers input_ratio (rotatecrop.rs:66-74)
        (w, h)
    }
    fn transform_reverse(&mut self, width: usize, height: usize) -> (usize, usize) {
        let mut pod = pod_rotatecrop(&self.op);
        let (mut w, mut h) = (0, 0);
        unsafe { ipb_rotatecrop_transform_reverse(&mut pod, width, height, &mut w, &mut h) };
        unpod_rotatecrop(&mut self.op, &pod); // ... and output_size (rotatecrop.rs:76-80)
        (w, h)
    }
    fn reset(&mut self) {
        let mut pod = pod_rotatecrop(&self.op);
        unsafe { ipb_rotatecrop_reset(&mut pod) };
        unpod_rotatecrop(&mut self.op, &pod);
    }
}

impl<'c> Gpu<'c, OpToLab> {
    pub fn run_device(&self, buf: &GpuBuffer) -> GpuBuffer {
        device_call!(self.ctx, ipb_tolab_run(&pod_tolab(&self.op), buf.raw))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpToLab> {
    fn name(&self) -> &str { "to_lab" }
    fn run(&self, _globals: &PipelineGlobals, buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        through_device(self.ctx, buf, |b| self.run_device(b))
    }
}

impl<'c> Gpu<'c, OpBaseCurve> {
    pub fn run_device(&self, buf: &GpuBuffer) -> GpuBuffer {
        device_call!(self.ctx, ipb_basecurve_run(&pod_basecurve(&self.op), buf.raw))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpBaseCurve> {
    fn name(&self) -> &str { "basecurve" }
    fn run(&self, _globals: &PipelineGlobals, buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        through_device(self.ctx, buf, |b| self.run_device(b))
    }
}

impl<'c> Gpu<'c, OpFromLab> {
    pub fn run_device(&self, buf: &GpuBuffer) -> GpuBuffer {
        device_call!(self.ctx, ipb_fromlab_run(buf.raw))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpFromLab> {
    fn name(&self) -> &str { "from_lab" }
    fn run(&self, _globals: &PipelineGlobals, buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        through_device(self.ctx, buf, |b| self.run_device(b))
    }
}

impl<'c> Gpu<'c, OpGamma> {
    pub fn run_device(&self, globals: &PipelineGlobals, buf: &GpuBuffer) -> GpuBuffer {
        device_call!(self.ctx, ipb_gamma_run(&pod_settings(&globals.settings), buf.raw))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpGamma> {
    fn name(&self) -> &str { "gamma" }
    fn run(&self, globals: &PipelineGlobals, buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        through_device(self.ctx, buf, |b| self.run_device(globals, b))
    }
}

impl<'c> Gpu<'c, OpTransform> {
    pub fn run_device(&self, buf: &GpuBuffer) -> GpuBuffer {
        device_call!(self.ctx, ipb_transform_run(&pod_transform(&self.op), buf.raw))
    }
}
impl<'a, 'c> ImageOp<'a> for Gpu<'c, OpTransform> {
    fn name(&self) -> &str { "transform" }
    fn run(&self, _globals: &PipelineGlobals, buf: Arc<OpBuffer>) -> Arc<OpBuffer> {
        through_device(self.ctx, buf, |b| self.run_device(b))
    }
    fn transform_forward(&mut self, width: usize, height: usize) -> (usize, usize) {
        let (mut w, mut h) = (0, 0);
        unsafe { ipb_transform_transform_forward(&pod_transform(&self.op), width, height, &mut w, &mut h) };
        (w, h)
    }
}

/// scaling_size / calculate_scale (src/scaling.rs:8-32) — host arithmetic, exposed for callers that size buffers
pub fn scaling_size(width: usize, height: usize, maxwidth: usize, maxheight: usize) -> (usize, usize) {
    let (mut w, mut h) = (0, 0);
    unsafe { ipb_scaling_size(width, height, maxwidth, maxheight, &mut w, &mut h) };
    (w, h)
}

// ------------------------------------------------------------------------------------------------ Pipeline hand-off

/// `PipelineCache` on the device (src/pipeline.rs:43, :257-260): an LRU of op outputs keyed by the cumulative hash of
/// settings and op parameters, `size` bytes of f32 data.
pub struct GpuCache {
    raw: *mut ipb_cache,
}
impl GpuCache {
    pub fn new(ctx: &GpuContext, size: usize) -> Self {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { ipb_cache_create(ctx.raw, size, &mut raw) });
        Self { raw }
    }
}
impl Drop for GpuCache {
    fn drop(&mut self) {
        unsafe { ipb_cache_destroy(self.raw) }
    }
}

/// The hand-off of `Pipeline::run` (src/pipeline.rs:311-375): reset, forward size walk, clamp to maxwidth / maxheight,
/// reverse walk to the demosaic size, then the eight ops — in one C call.  The library runs the chain as one fused
/// kernel when it can (u16 CFA source, rotatecrop a no-op) and op by op otherwise or with a cache; results are identical.
pub struct GpuPipeline<'c> {
    ctx: &'c GpuContext,
    raw: *mut ipb_pipeline,
    _source: SourceView, // the C side does not copy the pixels
}
impl<'c> GpuPipeline<'c> {
    /// Pipeline::new_from_source with the reference's own PipelineOps (which it derived from the image's metadata,
    /// src/pipeline.rs:166-179): parameters travel field by field.
    pub fn from_pipeline(ctx: &'c GpuContext, p: &Pipeline) -> Self {
        let source = source_view(&p.globals.image);
        let ops = pod_ops(&p.ops);
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { ipb_pipeline_create(ctx.raw, &source.pod, &ops, &mut raw) });
        let me = Self { ctx, raw, _source: source };
        me.sync_params(p);
        me
    }
    /// copy the (public, mutable) ops and settings of the reference pipeline before a run: callers edit them in place
    pub fn sync_params(&self, p: &Pipeline) {
        unsafe {
            *ipb_pipeline_ops(self.raw) = pod_ops(&p.ops);
            *ipb_pipeline_settings(self.raw) = pod_settings(&p.globals.settings);
        }
    }
    /// Pipeline::run(cache) -> Arc<OpBuffer> (3-channel f32, gamma-encoded unless settings.linear)
    pub fn run(&self, p: &mut Pipeline, cache: Option<&GpuCache>) -> Arc<OpBuffer> {
        self.sync_params(p);
        let mut out = ptr::null_mut();
        let rc = match cache {
            Some(c) => unsafe { ipb_pipeline_run_cached(self.raw, c.raw, &mut out) },
            None => unsafe { ipb_pipeline_run(self.raw, &mut out) },
        };
        self.ctx.check(rc);
        self.read_back_settings(p);
        Arc::new(GpuBuffer { raw: out }.download(self.ctx))
    }
    /// Pipeline::output_8bit (src/pipeline.rs:377-422), non-raw fast path included
    pub fn output_8bit(&self, p: &mut Pipeline, cache: Option<&GpuCache>) -> Result<SRGBImage, String> {
        self.sync_params(p);
        let (mut w, mut h) = (0, 0);
        self.ctx.check(unsafe { ipb_pipeline_output_size(self.raw, &mut w, &mut h) });
        let cap = std::cmp::max(w * h, self.src_pixels()) * 3;
        let mut data = vec![0u8; cap];
        let c = cache.map_or(ptr::null_mut(), |c| c.raw);
        let rc = unsafe { ipb_pipeline_output_8bit_cached(self.raw, c, data.as_mut_ptr(), cap, 0, &mut w, &mut h) };
        if rc != IPB_OK {
            return Err(last_error(self.ctx.raw));
        }
        self.read_back_settings(p);
        data.truncate(w * h * 3);
        Ok(SRGBImage { width: w, height: h, data })
    }
    /// Pipeline::output_16bit (src/pipeline.rs:424-469): linear, 16 bits per channel
    pub fn output_16bit(&self, p: &mut Pipeline, cache: Option<&GpuCache>) -> Result<SRGBImage16, String> {
        self.sync_params(p);
        let (mut w, mut h) = (0, 0);
        self.ctx.check(unsafe { ipb_pipeline_output_size(self.raw, &mut w, &mut h) });
        let cap = std::cmp::max(w * h, self.src_pixels()) * 3;
        let mut data = vec![0u16; cap];
        let c = cache.map_or(ptr::null_mut(), |c| c.raw);
        let rc = unsafe { ipb_pipeline_output_16bit_cached(self.raw, c, data.as_mut_ptr(), cap, 0, &mut w, &mut h) };
        if rc != IPB_OK {
            return Err(last_error(self.ctx.raw));
        }
        self.read_back_settings(p);
        data.truncate(w * h * 3);
        Ok(SRGBImage16 { width: w, height: h, data })
    }
    fn src_pixels(&self) -> usize {
        self._source.pod.width * self._source.pod.height
    }
    /// the reference leaves demosaic_width / demosaic_height / linear in globals.settings and rotatecrop's negotiated
    /// state in ops.rotatecrop after a run (src/pipeline.rs:337-338, :405, :452); mirror that
    fn read_back_settings(&self, p: &mut Pipeline) {
        unsafe {
            let s = &*ipb_pipeline_settings(self.raw);
            p.globals.settings.demosaic_width = s.demosaic_width;
            p.globals.settings.demosaic_height = s.demosaic_height;
            p.globals.settings.linear = s.linear != 0;
            unpod_rotatecrop(&mut p.ops.rotatecrop, &(*ipb_pipeline_ops(self.raw)).rotatecrop);
        }
    }
}
impl<'c> Drop for GpuPipeline<'c> {
    fn drop(&mut self) {
        unsafe { ipb_pipeline_destroy(self.raw) }
    }
}

/// What `Pipeline::run` becomes with the shim in place (the body of src/pipeline.rs:311-375 collapses to this):
///
/// ```ignore
/// pub fn run(&mut self, cache: Option<&PipelineCache>) -> Arc<OpBuffer> {
///     match &self.gpu {                      // Option<(GpuContext, GpuCache)> chosen at construction
///         Some((ctx, gcache)) => GpuPipeline::from_pipeline(ctx, self).run(self, cache.map(|_| gcache)),
///         None => self.run_cpu(cache),       // the existing Rayon path
///     }
/// }
/// ```
pub fn run_on_gpu(ctx: &GpuContext, p: &mut Pipeline, cache: Option<&GpuCache>) -> Arc<OpBuffer> {
    GpuPipeline::from_pipeline(ctx, p).run(p, cache)
}

// ------------------------------------------------------------------------------------------------ several GPUs
//
// One frame too large for one GPU (BASELINE config 5) is cut into row stripes, one process per GPU.  The reference has no
// counterpart; a host built on it would add this next to `Pipeline`.  Rank r owns the sensor rows of its stripe on its
// device; the rows its 3x3 stencil (or the resampler's windows) needs from the neighbours travel through
// `ipb_halo_exchange`, which resolves NCCL itself: the host only carries the 128-byte id from rank 0 to the others.

/// One rank's share of a frame: which output rows it produces and which source rows it holds (halo included).
#[derive(Clone, Copy, Debug)]
pub struct StripeLayout {
    pub out_row0: usize, pub out_row1: usize,   // output rows [out_row0, out_row1)
    pub src_row0: usize, pub src_row1: usize,   // source rows needed, halo included (ipb_stripe_plan)
    pub own_row0: usize, pub own_row1: usize,   // source rows this rank owns (the others arrive by exchange)
    pub out_width: usize, pub out_height: usize,
}

/// Even split of the output rows over `world` ranks; every rank computes the same table (host only, no device call).
pub fn plan_stripes(p: &Pipeline, width: usize, height: usize, world: usize) -> Vec<StripeLayout> {
    let (ops, settings) = (pod_ops(&p.ops), pod_settings(&p.globals.settings));
    let (mut s0, mut s1, mut ow, mut oh) = (0usize, 0usize, 0usize, 0usize);
    // the size of the result first (rows 0..0: no stripe, only the frame's output size)
    let rc = unsafe { ipb_stripe_plan(&ops, &settings, width, height, 0, 0, &mut s0, &mut s1, &mut ow, &mut oh) };
    assert_eq!(rc, IPB_OK, "only the fused raw CFA path with Normal orientation is cut into stripes");
    // balanced, interior boundaries on even rows (the Bayer period; any boundary is correct: kernels work in full-frame
    // coordinates)
    let bounds: Vec<usize> = (0..=world).map(|r| if r == world { oh } else { (oh * r / world + 1) / 2 * 2 }).collect();
    let mut lays: Vec<StripeLayout> = (0..world).map(|r| {
        let rc = unsafe { ipb_stripe_plan(&ops, &settings, width, height, bounds[r], bounds[r + 1], &mut s0, &mut s1, &mut ow, &mut oh) };
        assert_eq!(rc, IPB_OK);
        StripeLayout { out_row0: bounds[r], out_row1: bounds[r + 1], src_row0: s0, src_row1: s1, own_row0: s0, own_row1: s1,
                       out_width: ow, out_height: oh }
    }).collect();
    // ownership: the needed ranges overlap by the halo; the midpoint of each overlap separates the owners
    for r in 0..world.saturating_sub(1) {
        let cut = (lays[r].src_row1 + lays[r + 1].src_row0) / 2;
        lays[r].own_row1 = cut;
        lays[r + 1].own_row0 = cut;
    }
    if world > 0 { lays[0].own_row0 = 0; lays[world - 1].own_row1 = height; }
    lays
}

/// The byte ranges of a rank's stripe buffer that leave for / arrive from its neighbours (row_bytes = sensor width * 2).
pub fn halo_plan(lays: &[StripeLayout], rank: usize, row_bytes: usize) -> ipb_halo {
    let me = &lays[rank];
    let off = |row: usize| (row - me.src_row0) * row_bytes;
    let mut h = ipb_halo { send_up_off: 0, send_up_bytes: 0, recv_up_off: 0, recv_up_bytes: 0,
                           send_down_off: 0, send_down_bytes: 0, recv_down_off: 0, recv_down_bytes: 0 };
    if rank > 0 {
        let up = &lays[rank - 1];
        // what the upper neighbour needs of my rows, what I need of its rows
        h.send_up_off = off(me.own_row0); h.send_up_bytes = (up.src_row1 - me.own_row0) * row_bytes;
        h.recv_up_off = off(me.src_row0); h.recv_up_bytes = (me.own_row0 - me.src_row0) * row_bytes;
    }
    if rank + 1 < lays.len() {
        let down = &lays[rank + 1];
        h.send_down_off = off(down.src_row0); h.send_down_bytes = (me.own_row1 - down.src_row0) * row_bytes;
        h.recv_down_off = off(me.own_row1); h.recv_down_bytes = (me.src_row1 - me.own_row1) * row_bytes;
    }
    h
}

/// The communicator of the stripe neighbours (one per process).  `id`: 128 bytes drawn by rank 0 with
/// `StripeComm::unique_id()` and handed to every rank by the host program (MPI, a file, a socket).
pub struct StripeComm { raw: *mut ipb_comm }
impl StripeComm {
    pub fn unique_id() -> Result<[u8; 128], String> {
        let mut id = [0u8; 128];
        if unsafe { ipb_comm_unique_id(id.as_mut_ptr()) } != IPB_OK { return Err(comm_error(ptr::null())); }
        Ok(id)
    }
    pub fn new(device: i32, stream: *mut std::os::raw::c_void, id: &[u8; 128], rank: usize, world: usize) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        if unsafe { ipb_comm_create(device, stream, id.as_ptr(), rank as c_int, world as c_int, &mut raw) } != IPB_OK {
            return Err(comm_error(ptr::null()));
        }
        Ok(Self { raw })
    }
    /// Halo rows of every buffer in `bufs` (the frames in flight; device addresses of the stripes' first source row):
    /// one packed message per neighbour and direction, enqueued on the communicator's stream.
    pub fn exchange(&self, bufs: &[*const std::os::raw::c_void], halo: &ipb_halo) -> Result<(), String> {
        if unsafe { ipb_halo_exchange(self.raw, bufs.as_ptr(), bufs.len(), halo) } != IPB_OK { return Err(comm_error(self.raw)); }
        Ok(())
    }
}
impl Drop for StripeComm {
    fn drop(&mut self) { unsafe { ipb_comm_destroy(self.raw) } }
}
fn comm_error(c: *const ipb_comm) -> String {
    unsafe { std::ffi::CStr::from_ptr(ipb_comm_last_error(c)) }.to_string_lossy().into_owned()
}

impl<'c> GpuPipeline<'c> {
    /// This rank's stripes of `nframes` frames in flight: `rows` is the device address of the first stripe's first source
    /// row (`lay.src_row0`), the stripes follow each other (`lay.src_row1 - lay.src_row0` rows apart), their halo rows
    /// already exchanged; `dst` (device) receives the `nframes` results one after the other.  One launch for all of them
    /// where the speculative kernel applies (ipb_pipeline_output_8bit_batch on the stripe source).
    pub fn output_8bit_stripes(&self, rows: *const std::os::raw::c_void, sensor_width: usize, full_height: usize, lay: &StripeLayout,
                               nframes: usize, dst: *mut u8, dst_capacity: usize) -> Result<(usize, usize), String> {
        let nrows = lay.src_row1 - lay.src_row0;
        let src = ipb_source { kind: IPB_SRC_RAW_U16, width: sensor_width, height: nrows, cpp: 1, data: rows, on_device: 1 };
        let st = ipb_stripe { full_height, src_row0: lay.src_row0, out_row0: lay.out_row0, out_row1: lay.out_row1 };
        if unsafe { ipb_pipeline_set_stripe_source(self.raw, &src, &st) } != IPB_OK { return Err(last_error(self.ctx.raw)); }
        let (mut w, mut h) = (0usize, 0usize);
        let stride = (lay.out_row1 - lay.out_row0) * lay.out_width * 3;
        let rc = unsafe { ipb_pipeline_output_8bit_batch(self.raw, nframes, nrows, dst, stride, dst_capacity, &mut w, &mut h) };
        if rc != IPB_OK { return Err(last_error(self.ctx.raw)); }
        Ok((w, h))
    }
}

/// A step of the sharded pipeline on one rank, in the order the calls have to be made:
///
/// ```ignore
/// let lays = plan_stripes(&pipeline, width, height, world);          // every rank, host only
/// let me = lays[rank];
/// let halo = halo_plan(&lays, rank, width * 2);
/// let comm = StripeComm::new(device, stream, &id, rank, world)?;      // id: from rank 0, over the host's own channel
/// // ... the rank's own rows of each frame in flight are on the device (bufs[k] + byte offset of me.own_row0) ...
/// comm.exchange(&bufs, &halo)?;                                       // neighbours' rows arrive in place
/// gpu_pipeline.output_8bit_stripes(bufs[0], width, height, &me, bufs.len(), dst, cap)?;   // stripes -> sRGB bytes
/// ctx.synchronize();
/// ```
pub const SHARDED_STEP: () = ();
