// ipb_internal.h — interface between the C++ host layer (ipb_host.cu) and the kernel launchers
// (ipb_ops.cu, ipb_fused.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "ipb_device.cuh"

namespace ipb {

// rawloader::CFA colour table, tiled to 48x48 like the reference (demosaic.rs:77-95 indexes [row%48][col%48]).
struct CfaDev {
  int width, height;       // pattern period (2x2 Bayer, 6x6 X-Trans, 2x8, 12x12); 0 = no CFA
  uint8_t pat[48 * 48];
};

// transform_buffer corner points (scaling.rs:51-63)
struct XformGeom {
  long tl[2], tr[2], bl[2];
  size_t width, height;    // source
  size_t nwidth, nheight;  // destination
  size_t components;
};

enum FusedOut { kOutF32 = 0, kOutU8 = 1, kOutU16 = 2 };

// Arguments of the fused raw->sRGB kernels.  Rows are in full-frame coordinates so that row stripes
// (multi-GPU / chunked host transfers) see the right CFA phase and image borders.
struct FusedArgs {
  const uint16_t *raw;     // first available source row (full-frame row src_row0), un-cropped sensor width
  size_t raw_pitch;        // elements per source row (owidth)
  size_t src_row0;         // full-frame (un-cropped) index of raw's first row
  size_t src_rows;         // rows available in raw
  size_t crop_x, crop_y;   // gofloat crop origin (gofloat.rs:74-82)
  size_t width, height;    // full cropped frame size (== demosaic input size)
  size_t out_row0, out_row1;  // output rows to produce (output-frame coordinates)
  size_t out_width, out_height;  // full output frame (== width,height for the full-res kernel)
  void *out;               // row out_row0 of the output frame, 3 interleaved channels
  int out_kind;            // FusedOut
  float black, range, range_rc;
  int exact_rc;
  const float2 *lut_lab;   // device tables {v, dv}
  const float2 *lut_gamma;
  const float2 *lut_gamma8;  // 8-bit output: {threshold, base} per table segment (ipb_host.cu build_gamma8)
  const float *cbrt_tab;     // full-res kernel: host cbrtf of every float in (1.0, 1.5] (ipb_host.cu ensure_cbrt_table)
  int use_tma;             // full-res kernel: stage tiles with TMA (needs 16B-aligned base and pitch)
  // a batch of frames with this geometry in one launch (k_spec8 only): frame k's source rows start batch_src_rows * k
  // rows after raw, its output batch_out_bytes * k bytes after out.  batch_n <= 1: a single frame.
  size_t batch_n, batch_src_rows, batch_out_bytes;
};

// ---- ipb_ops.cu
cudaError_t launch_gofloat_raw(cudaStream_t s, int is_f32, const void *src, size_t total_elems, size_t owidth,
                               size_t x, size_t y, size_t width, size_t height, size_t cpp, int mode,
                               const float mins[4], const float ranges[4], int rc_exact, float *out);
bool gofloat_rows_cover(size_t total, size_t owidth, size_t x, size_t y, size_t width, size_t height, size_t cpp);
cudaError_t launch_gofloat_other(cudaStream_t s, int is16, const void *src, size_t owidth, size_t x, size_t y,
                                 size_t width, size_t height, const float2 *lut_rev, float *out);
cudaError_t launch_demosaic_full(cudaStream_t s, const CfaDev &cfa, const float *in, size_t w, size_t h, float *out);
cudaError_t launch_transform_f32(cudaStream_t s, const XformGeom &g, const CfaDev *cfa, const float *src, float *out);
cudaError_t launch_transform_u8(cudaStream_t s, const XformGeom &g, const uint8_t *src, uint8_t *out);
cudaError_t launch_transform_u16(cudaStream_t s, const XformGeom &g, const uint16_t *src, uint16_t *out);
// curve != 0: P.sp is applied to L before the store (to_lab + basecurve in one pass); 1 = the reference's binary search,
// 2 = the counting form for finite, strictly increasing knots with finite coefficients.  cbrt_tab may be null.
cudaError_t launch_tolab(cudaStream_t s, const ColorParams &P, const float2 *lut_lab, const float *cbrt_tab, int curve,
                         const float *in, size_t npix, float *out);
cudaError_t launch_basecurve(cudaStream_t s, const SplineDev &sp, const float *in, size_t npix, float *out);
// lut_gamma != null: OpGamma is applied before the store (from_lab + gamma in one pass)
cudaError_t launch_fromlab(cudaStream_t s, const ColorParams &P, const float2 *lut_gamma, const float *in, size_t npix,
                           float *out);
cudaError_t launch_gamma(cudaStream_t s, const float2 *lut_gamma, const float *in, size_t nelem, float *out);
cudaError_t launch_pack8(cudaStream_t s, const float *in, size_t nelem, uint8_t *out);
cudaError_t launch_pack16(cudaStream_t s, const float *in, size_t nelem, uint16_t *out);
cudaError_t launch_rotate(cudaStream_t s, const float *in, size_t w, size_t h, int transpose, int flip_x, int flip_y,
                          float *out);
cudaError_t launch_synth(cudaStream_t s, uint64_t seed, size_t width, size_t row0, size_t rows, uint16_t *out);
cudaError_t launch_spline_eval(cudaStream_t s, const SplineDev &sp, const float *in, size_t n, float *out);
cudaError_t launch_rgb16_to_8(cudaStream_t s, const uint16_t *in, size_t n, uint8_t *out);
cudaError_t launch_rgb8_to_16(cudaStream_t s, const uint8_t *in, size_t n, uint16_t *out);

// ---- ipb_lanczos.cu (extension: Lanczos-a separable resampler, horizontal then vertical pass; two launches)
cudaError_t launch_lanczos(cudaStream_t s, const float *in, size_t W, size_t H, size_t C, size_t nw, size_t nh,
                           const int *sx, const int *cx, const float *wx, int kx, size_t max_span_floats, const int *sy,
                           const int *cy, const float *wy, int ky, float *mid, float *out);

// ---- ipb_fused.cu
// full-resolution CFA -> RGB (gofloat + demosaic::full + colour chain + pack) in one kernel
cudaError_t launch_fused_full(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                              int sm_count);
// scaled_demosaic (scaling.rs:132-145) fused with gofloat, the colour chain and the pack
cudaError_t launch_fused_scaled(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                                int sm_count);
const char *fused_last_error();
// self-test of the 8-bit gamma threshold table: every f32 in [0,1] (and a few outside) through gamma8() vs
// output8bit(gamma_elem()); *mismatches (device) receives the count
cudaError_t launch_gamma8_selftest(cudaStream_t s, const float2 *lut_gamma, const float2 *lut_gamma8,
                                   unsigned long long *mismatches);
// gamma + 8-bit pack of arbitrary floats through the threshold table (tests)
cudaError_t launch_gamma8_pack(cudaStream_t s, const float2 *lut_gamma8, const float *in, size_t n, uint8_t *out);

}  // namespace ipb
