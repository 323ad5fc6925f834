// ffi.rs — GENERATED from include/ipb200.h by tools/gen_ffi_rs.py (do not edit; tests/test_shim_rs.py checks it).
// The `extern "C"` block the reference (pedrocr/imagepipe) binds libipb200.so with: one declaration per C
// entry point, `#[repr(C)]` twins of the parameter structs (which mirror the reference's serde structs field
// by field: see the comments of ipb200.h), the status and enum constants.  NOT compiled here: this image has
// no Rust toolchain.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_long, c_ulonglong, c_void};

pub const IPB_VERSION: usize = 100;
pub const IPB_MAX_CURVE_POINTS: usize = 32;
pub const IPB_COMM_ID_BYTES: usize = 128;
pub const IPB_OK: c_int = 0;
pub const IPB_ERR_INVALID: c_int = 1;
pub const IPB_ERR_BAD_COLORS: c_int = 2;
pub const IPB_ERR_BAD_CFA: c_int = 3;
pub const IPB_ERR_CUDA: c_int = 4;
pub const IPB_ERR_UNSUPPORTED: c_int = 5;
pub const IPB_ERR_NOMEM: c_int = 6;
pub const IPB_ROT_NORMAL: c_int = 0;
pub const IPB_ROT_90: c_int = 1;
pub const IPB_ROT_180: c_int = 2;
pub const IPB_ROT_270: c_int = 3;
pub const IPB_SRC_RAW_U16: c_int = 0;
pub const IPB_SRC_RAW_F32: c_int = 1;
pub const IPB_SRC_RGB8: c_int = 2;
pub const IPB_SRC_RGB16: c_int = 3;

#[repr(C)] pub struct ipb_ctx { _private: [u8; 0] }
#[repr(C)] pub struct ipb_buffer { _private: [u8; 0] }
#[repr(C)] pub struct ipb_pipeline { _private: [u8; 0] }
#[repr(C)] pub struct ipb_cache { _private: [u8; 0] }
#[repr(C)] pub struct ipb_comm { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_gofloat {
    pub crop_top: usize,
    pub crop_right: usize,
    pub crop_bottom: usize,
    pub crop_left: usize,
    pub is_cfa: c_int,
    pub blacklevels: [f32; 4],
    pub whitelevels: [f32; 4],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_demosaic {
    pub cfa: [c_char; 148],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_rotatecrop {
    pub crop_top: f32,
    pub crop_right: f32,
    pub crop_bottom: f32,
    pub crop_left: f32,
    pub rotation: f32,
    pub input_ratio: f32,
    pub has_output_size: c_int,
    pub output_width: usize,
    pub output_height: usize,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_tolab {
    pub cam_to_xyz: [[f32; 4]; 3],
    pub cam_to_xyz_normalized: [[f32; 4]; 3],
    pub xyz_to_cam: [[f32; 3]; 4],
    pub wb_coeffs: [f32; 4],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_basecurve {
    pub exposure: f32,
    pub npoints: usize,
    pub points: [[f32; 2]; 32],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_transform {
    pub rotation: c_int,
    pub fliph: c_int,
    pub flipv: c_int,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_settings {
    pub maxwidth: usize,
    pub maxheight: usize,
    pub demosaic_width: usize,
    pub demosaic_height: usize,
    pub linear: c_int,
    pub use_fastpath: c_int,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_source {
    pub kind: c_int,
    pub width: usize,
    pub height: usize,
    pub cpp: usize,
    pub data: *const c_void,
    pub on_device: c_int,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_ops {
    pub gofloat: ipb_gofloat,
    pub demosaic: ipb_demosaic,
    pub rotatecrop: ipb_rotatecrop,
    pub tolab: ipb_tolab,
    pub basecurve: ipb_basecurve,
    pub transform: ipb_transform,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_stripe {
    pub full_height: usize,
    pub src_row0: usize,
    pub out_row0: usize,
    pub out_row1: usize,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ipb_halo {
    pub send_up_off: usize,
    pub send_up_bytes: usize,
    pub recv_up_off: usize,
    pub recv_up_bytes: usize,
    pub send_down_off: usize,
    pub send_down_bytes: usize,
    pub recv_down_off: usize,
    pub recv_down_bytes: usize,
}

#[link(name = "ipb200")]
extern "C" {
    pub fn ipb_version() -> c_int;
    pub fn ipb_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut ipb_ctx) -> c_int;
    pub fn ipb_ctx_destroy(ctx: *mut ipb_ctx);
    pub fn ipb_ctx_set_stream(ctx: *mut ipb_ctx, stream: *mut c_void) -> c_int;
    pub fn ipb_ctx_synchronize(ctx: *mut ipb_ctx) -> c_int;
    pub fn ipb_last_error(ctx: *const ipb_ctx) -> *const c_char;
    pub fn ipb_ctx_launch_count(ctx: *const ipb_ctx) -> c_ulonglong;
    pub fn ipb_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn ipb_host_free(p: *mut c_void);
    pub fn ipb_device_alloc(ctx: *mut ipb_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn ipb_device_free(ctx: *mut ipb_ctx, dptr: *mut c_void) -> c_int;
    pub fn ipb_device_upload(ctx: *mut ipb_ctx, dptr: *mut c_void, host: *const c_void, bytes: usize) -> c_int;
    pub fn ipb_device_download(ctx: *mut ipb_ctx, host: *mut c_void, dptr: *const c_void, bytes: usize) -> c_int;
    pub fn ipb_buffer_new(ctx: *mut ipb_ctx, width: usize, height: usize, colors: usize, monochrome: c_int, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_buffer_upload(ctx: *mut ipb_ctx, width: usize, height: usize, colors: usize, monochrome: c_int, host: *const f32, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_buffer_wrap(ctx: *mut ipb_ctx, width: usize, height: usize, colors: usize, monochrome: c_int, dptr: *mut c_void, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_buffer_download(ctx: *mut ipb_ctx, buf: *const ipb_buffer, host: *mut f32) -> c_int;
    pub fn ipb_buffer_retain(buf: *mut ipb_buffer);
    pub fn ipb_buffer_release(buf: *mut ipb_buffer);
    pub fn ipb_buffer_width(buf: *const ipb_buffer) -> usize;
    pub fn ipb_buffer_height(buf: *const ipb_buffer) -> usize;
    pub fn ipb_buffer_colors(buf: *const ipb_buffer) -> usize;
    pub fn ipb_buffer_monochrome(buf: *const ipb_buffer) -> c_int;
    pub fn ipb_buffer_device_ptr(buf: *const ipb_buffer) -> *mut c_void;
    pub fn ipb_gofloat_run(ctx: *mut ipb_ctx, op: *const ipb_gofloat, image: *const ipb_source, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_demosaic_run(ctx: *mut ipb_ctx, op: *const ipb_demosaic, settings: *const ipb_settings, r#in: *mut ipb_buffer, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_rotatecrop_run(ctx: *mut ipb_ctx, op: *const ipb_rotatecrop, r#in: *mut ipb_buffer, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_tolab_run(ctx: *mut ipb_ctx, op: *const ipb_tolab, r#in: *mut ipb_buffer, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_basecurve_run(ctx: *mut ipb_ctx, op: *const ipb_basecurve, r#in: *mut ipb_buffer, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_fromlab_run(ctx: *mut ipb_ctx, r#in: *mut ipb_buffer, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_gamma_run(ctx: *mut ipb_ctx, settings: *const ipb_settings, r#in: *mut ipb_buffer, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_transform_run(ctx: *mut ipb_ctx, op: *const ipb_transform, r#in: *mut ipb_buffer, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_gofloat_transform_forward(op: *const ipb_gofloat, w: usize, h: usize, ow: *mut usize, oh: *mut usize);
    pub fn ipb_rotatecrop_transform_forward(op: *mut ipb_rotatecrop, w: usize, h: usize, ow: *mut usize, oh: *mut usize);
    pub fn ipb_rotatecrop_transform_reverse(op: *mut ipb_rotatecrop, w: usize, h: usize, ow: *mut usize, oh: *mut usize);
    pub fn ipb_rotatecrop_reset(op: *mut ipb_rotatecrop);
    pub fn ipb_transform_transform_forward(op: *const ipb_transform, w: usize, h: usize, ow: *mut usize, oh: *mut usize);
    pub fn ipb_scaling_size(w: usize, h: usize, maxw: usize, maxh: usize, ow: *mut usize, oh: *mut usize);
    pub fn ipb_calculate_scale(w: usize, h: usize, maxw: usize, maxh: usize) -> f32;
    pub fn ipb_spline_eval(ctx: *mut ipb_ctx, op: *const ipb_basecurve, r#in: *const f32, out: *mut f32, n: usize) -> c_int;
    pub fn ipb_pack_8bit(ctx: *mut ipb_ctx, r#in: *const ipb_buffer, dst: *mut u8, dst_on_device: c_int) -> c_int;
    pub fn ipb_pack_16bit(ctx: *mut ipb_ctx, r#in: *const ipb_buffer, dst: *mut u16, dst_on_device: c_int) -> c_int;
    pub fn ipb_scale_down_srgb(ctx: *mut ipb_ctx, src: *const u8, w: usize, h: usize, nw: usize, nh: usize, dst: *mut u8, on_device: c_int) -> c_int;
    pub fn ipb_scale_down_srgb16(ctx: *mut ipb_ctx, src: *const u16, w: usize, h: usize, nw: usize, nh: usize, dst: *mut u16, on_device: c_int) -> c_int;
    pub fn ipb_lanczos_resize(ctx: *mut ipb_ctx, r#in: *mut ipb_buffer, nwidth: usize, nheight: usize, a: c_int, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_ops_default(ops: *mut ipb_ops, image: *const ipb_source);
    pub fn ipb_pipeline_create(ctx: *mut ipb_ctx, image: *const ipb_source, ops: *const ipb_ops, out: *mut *mut ipb_pipeline) -> c_int;
    pub fn ipb_pipeline_destroy(p: *mut ipb_pipeline);
    pub fn ipb_pipeline_ops(p: *mut ipb_pipeline) -> *mut ipb_ops;
    pub fn ipb_pipeline_settings(p: *mut ipb_pipeline) -> *mut ipb_settings;
    pub fn ipb_pipeline_set_source(p: *mut ipb_pipeline, image: *const ipb_source) -> c_int;
    pub fn ipb_pipeline_set_fused(p: *mut ipb_pipeline, fused: c_int) -> c_int;
    pub fn ipb_pipeline_set_tma(p: *mut ipb_pipeline, use_tma: c_int) -> c_int;
    pub fn ipb_pipeline_set_speculative(p: *mut ipb_pipeline, on: c_int) -> c_int;
    pub fn ipb_ctx_set_spec(ctx: *mut ipb_ctx, delta: f32, threads: c_int) -> c_int;
    pub fn ipb_ctx_spec_stats(ctx: *mut ipb_ctx, out: *mut c_ulonglong, reset: c_int) -> c_int;
    pub fn ipb_pipeline_spec_probe(p: *mut ipb_pipeline, max_dev: *mut f32, mean_dev: *mut f64, delta: *mut f32) -> c_int;
    pub fn ipb_spec_bound(ops: *const ipb_ops, mufu_rel_err: f32, delta: *mut f32) -> c_int;
    pub fn ipb_spec_tables(ops: *const ipb_ops, mufu_rel_err: f32, delta_override: f32, g8a: *mut u32, thresholds: *mut f32, one: *mut f32, wmul: *mut u32, amb_t: *mut u32, delta: *mut f32) -> c_int;
    pub fn ipb_scaled_division_check(width: usize, height: usize, nwidth: usize, nheight: usize) -> c_int;
    pub fn ipb_pipeline_set_band_mb(p: *mut ipb_pipeline, megabytes: c_int) -> c_int;
    pub fn ipb_pipeline_output_size(p: *mut ipb_pipeline, width: *mut usize, height: *mut usize) -> c_int;
    pub fn ipb_pipeline_run(p: *mut ipb_pipeline, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_cache_create(ctx: *mut ipb_ctx, max_bytes: usize, out: *mut *mut ipb_cache) -> c_int;
    pub fn ipb_cache_destroy(cache: *mut ipb_cache);
    pub fn ipb_cache_clear(cache: *mut ipb_cache);
    pub fn ipb_cache_bytes(cache: *const ipb_cache) -> usize;
    pub fn ipb_cache_entries(cache: *const ipb_cache) -> usize;
    pub fn ipb_pipeline_run_cached(p: *mut ipb_pipeline, cache: *mut ipb_cache, out: *mut *mut ipb_buffer) -> c_int;
    pub fn ipb_pipeline_last_run_info(p: *const ipb_pipeline, startpos: *mut c_int, ops_run: *mut c_int);
    pub fn ipb_pipeline_output_8bit(p: *mut ipb_pipeline, dst: *mut u8, dst_capacity: usize, dst_on_device: c_int, width: *mut usize, height: *mut usize) -> c_int;
    pub fn ipb_pipeline_output_16bit(p: *mut ipb_pipeline, dst: *mut u16, dst_capacity: usize, dst_on_device: c_int, width: *mut usize, height: *mut usize) -> c_int;
    pub fn ipb_pipeline_output_8bit_cached(p: *mut ipb_pipeline, cache: *mut ipb_cache, dst: *mut u8, dst_capacity: usize, dst_on_device: c_int, width: *mut usize, height: *mut usize) -> c_int;
    pub fn ipb_pipeline_output_16bit_cached(p: *mut ipb_pipeline, cache: *mut ipb_cache, dst: *mut u16, dst_capacity: usize, dst_on_device: c_int, width: *mut usize, height: *mut usize) -> c_int;
    pub fn ipb_pipeline_stripe_rows(p: *mut ipb_pipeline, out_row0: usize, out_row1: usize, src_row0: *mut usize, src_row1: *mut usize) -> c_int;
    pub fn ipb_stripe_plan(ops: *const ipb_ops, settings: *const ipb_settings, width: usize, height: usize, out_row0: usize, out_row1: usize, src_row0: *mut usize, src_row1: *mut usize, out_width: *mut usize, out_height: *mut usize) -> c_int;
    pub fn ipb_pipeline_set_stripe_source(p: *mut ipb_pipeline, rows: *const ipb_source, stripe: *const ipb_stripe) -> c_int;
    pub fn ipb_pipeline_output_8bit_stripe(p: *mut ipb_pipeline, dst: *mut u8, dst_capacity: usize, dst_on_device: c_int, width: *mut usize, rows: *mut usize) -> c_int;
    pub fn ipb_pipeline_output_8bit_batch(p: *mut ipb_pipeline, nframes: usize, src_stride_rows: usize, dst: *mut u8, dst_stride_bytes: usize, dst_capacity: usize, width: *mut usize, rows: *mut usize) -> c_int;
    pub fn ipb_comm_unique_id(id: *mut u8) -> c_int;
    pub fn ipb_comm_create(device: c_int, stream: *mut c_void, id: *const u8, rank: c_int, nranks: c_int, out: *mut *mut ipb_comm) -> c_int;
    pub fn ipb_comm_destroy(comm: *mut ipb_comm);
    pub fn ipb_comm_rank(comm: *const ipb_comm) -> c_int;
    pub fn ipb_comm_size(comm: *const ipb_comm) -> c_int;
    pub fn ipb_comm_nccl_version(version: *mut c_int) -> c_int;
    pub fn ipb_comm_last_error(comm: *const ipb_comm) -> *const c_char;
    pub fn ipb_halo_exchange(comm: *mut ipb_comm, bufs: *const *const c_void, nbufs: usize, halo: *const ipb_halo) -> c_int;
    pub fn ipb_selftest_gamma8(ctx: *mut ipb_ctx, mismatches: *mut c_ulonglong) -> c_int;
    pub fn ipb_gamma_pack_8bit(ctx: *mut ipb_ctx, r#in: *const f32, n: usize, out: *mut u8) -> c_int;
    pub fn ipb_synth_cfa_u16(ctx: *mut ipb_ctx, seed: u64, width: usize, row0: usize, rows: usize, dptr: *mut u16) -> c_int;
}
