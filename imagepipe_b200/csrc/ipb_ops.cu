// ipb_ops.cu — one CUDA kernel per ImageOp of the reference (the "unfused" path): each op reads one
// device OpBuffer and writes a new one, exactly like the reference's per-op passes.  These kernels are
// the per-op entry points of the C ABI (ipb_*_run) and the building blocks for chains the fused kernels
// do not cover.  They are HBM-bound elementwise / stencil passes: one thread per pixel (or element),
// consecutive threads on consecutive addresses, grid sized to cover the buffer.
#include "ipb_internal.h"

namespace ipb {

static inline unsigned grid_for(size_t n, unsigned block) {
  size_t g = (n + block - 1) / block;
  return (unsigned)(g ? g : 1);
}

// ------------------------------------------------------------------ K1 gofloat (gofloat.rs:84-201)

// true when every sample the CFA / plain branch reads lies inside the raster (then no output element stays at its
// initial 0.0 and the buffer needs no zero fill)
bool gofloat_rows_cover(size_t total, size_t owidth, size_t x, size_t y, size_t width, size_t height, size_t cpp) {
  if (width == 0 || height == 0) return true;
  return owidth * (height - 1 + y) + x + width * cpp <= total;
}

template <typename T>
__global__ void k_gofloat_raw(const T *__restrict__ src, size_t total, size_t owidth, size_t x, size_t y,
                              size_t width, size_t height, size_t cpp, int mode, float m0, float m1, float m2,
                              float r0, float r1, float r2, float *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (mode == 2) {  // CFA / plain: cpp channels, level index 0 only (gofloat.rs:122-130)
    size_t linelen = width * cpp;
    if (idx >= linelen * height) return;
    size_t row = idx / linelen, c = idx - row * linelen;
    size_t off = owidth * (row + y) + x + c;
    if (off < total) out[idx] = fminf(__fdiv_rn((float)src[off] - m0, r0), 1.0f);
    return;
  }
  if (idx >= width * height) return;
  size_t row = idx / width, col = idx - row * width;
  float4 o;
  if (mode == 0) {  // monochrome -> RGB (gofloat.rs:97-109)
    float v = fminf(__fdiv_rn((float)src[owidth * (row + y) + x + col] - m0, r0), 1.0f);
    o = make_float4(v, v, v, 0.0f);
  } else {  // 3 cpp -> four channel (gofloat.rs:110-121)
    const T *p = src + (owidth * (row + y) + x + col) * 3;
    o.x = fminf(__fdiv_rn((float)p[0] - m0, r0), 1.0f);
    o.y = fminf(__fdiv_rn((float)p[1] - m1, r1), 1.0f);
    o.z = fminf(__fdiv_rn((float)p[2] - m2, r2), 1.0f);
    o.w = 0.0f;
  }
  reinterpret_cast<float4 *>(out)[idx] = o;
}

// mode 2 (CFA / plain, gofloat.rs:122-130) without 64-bit index arithmetic: blockIdx.x = output row, blockIdx.y =
// 256-element chunk of the row; 2-byte loads and 4-byte stores on consecutive addresses.  The caller has checked that
// every source offset is inside the raster (otherwise the general kernel runs on a zeroed buffer).
template <typename T>
__global__ void __launch_bounds__(256)
k_gofloat_rows(const T *__restrict__ src, size_t owidth, size_t x, size_t y, unsigned linelen, float m0, float r0,
               float *__restrict__ out) {
  const unsigned c = blockIdx.y * 256u + threadIdx.x;
  if (c >= linelen) return;
  const size_t row = blockIdx.x;
  const float v = (float)__ldg(src + owidth * (row + y) + x + c);
  out[row * linelen + c] = fminf(__fdiv_rn(v - m0, r0), 1.0f);
}

// The same for u16 sources whose rows start on 16-byte boundaries: eight samples per thread (one 16-byte load, two
// 16-byte stores), u16 -> f32 through the exact 2^23 + v identity, and — when the host has verified it for every
// possible sample (golevel_rc_exact) — the three-instruction reciprocal form of the division, else IEEE division.
template <bool RC>
__global__ void __launch_bounds__(256)
k_gofloat_rows8(const uint16_t *__restrict__ src, size_t owidth, size_t x, size_t y, unsigned linelen, float m0, float r0,
                float rc, float *__restrict__ out) {
  const unsigned c = (blockIdx.y * 256u + threadIdx.x) * 8u;
  if (c >= linelen) return;
  const size_t row = blockIdx.x;
  const uint4 pk = __ldg(reinterpret_cast<const uint4 *>(src + owidth * (row + y) + x + c));
  const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float lo = __uint_as_float(0x4B000000u | (w[k] & 0xffffu)) - 8388608.0f;
    const float hi = __uint_as_float(0x4B000000u | (w[k] >> 16)) - 8388608.0f;
    v[2 * k] = fminf(RC ? div_rc(lo - m0, r0, rc) : __fdiv_rn(lo - m0, r0), 1.0f);
    v[2 * k + 1] = fminf(RC ? div_rc(hi - m0, r0, rc) : __fdiv_rn(hi - m0, r0), 1.0f);
  }
  float4 *o = reinterpret_cast<float4 *>(out + row * linelen + c);
  o[0] = make_float4(v[0], v[1], v[2], v[3]);
  o[1] = make_float4(v[4], v[5], v[6], v[7]);
}

cudaError_t launch_gofloat_raw(cudaStream_t s, int is_f32, const void *src, size_t total, size_t owidth, size_t x,
                               size_t y, size_t width, size_t height, size_t cpp, int mode, const float mins[4],
                               const float ranges[4], int rc_exact, float *out) {
  size_t n = mode == 2 ? width * cpp * height : width * height;
  if (n == 0) return cudaSuccess;
  const size_t linelen = width * cpp;
  if (mode == 2 && gofloat_rows_cover(total, owidth, x, y, width, height, cpp) && (linelen + 255) / 256 <= 65535 &&
      height < (1ull << 31)) {
    const bool vec8 = !is_f32 && (linelen % 8) == 0 && (owidth % 8) == 0 && (x % 8) == 0 &&
                      (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (vec8) {
      dim3 grid((unsigned)height, (unsigned)((linelen / 8 + 255) / 256));
      if (rc_exact)
        k_gofloat_rows8<true><<<grid, 256, 0, s>>>((const uint16_t *)src, owidth, x, y, (unsigned)linelen, mins[0], ranges[0],
                                                   1.0f / ranges[0], out);
      else
        k_gofloat_rows8<false><<<grid, 256, 0, s>>>((const uint16_t *)src, owidth, x, y, (unsigned)linelen, mins[0], ranges[0],
                                                    0.0f, out);
      return cudaGetLastError();
    }
    dim3 grid((unsigned)height, (unsigned)((linelen + 255) / 256));
    if (is_f32)
      k_gofloat_rows<float><<<grid, 256, 0, s>>>((const float *)src, owidth, x, y, (unsigned)linelen, mins[0], ranges[0], out);
    else
      k_gofloat_rows<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)src, owidth, x, y, (unsigned)linelen, mins[0],
                                                    ranges[0], out);
    return cudaGetLastError();
  }
  if (is_f32)
    k_gofloat_raw<float><<<grid_for(n, 256), 256, 0, s>>>((const float *)src, total, owidth, x, y, width, height, cpp,
                                                          mode, mins[0], mins[1], mins[2], ranges[0], ranges[1],
                                                          ranges[2], out);
  else
    k_gofloat_raw<uint16_t><<<grid_for(n, 256), 256, 0, s>>>((const uint16_t *)src, total, owidth, x, y, width,
                                                             height, cpp, mode, mins[0], mins[1], mins[2], ranges[0],
                                                             ranges[1], ranges[2], out);
  return cudaGetLastError();
}

// run_other (gofloat.rs:171-201): RGB8 through input8bit + expand_srgb_gamma, RGB16 through input16bit
template <typename T>
__global__ void k_gofloat_other(const T *__restrict__ src, size_t owidth, size_t x, size_t y, size_t width,
                                size_t height, const float2 *__restrict__ lut_rev, float *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= width * height) return;
  size_t row = idx / width, col = idx - row * width;
  const T *p = src + (owidth * (row + y) + x + col) * 3;
  float4 o;
  if (sizeof(T) == 1) {
    LutGlobal lut{lut_rev};
    o.x = lut_lerp(lut, __fdiv_rn((float)p[0], 255.0f));
    o.y = lut_lerp(lut, __fdiv_rn((float)p[1], 255.0f));
    o.z = lut_lerp(lut, __fdiv_rn((float)p[2], 255.0f));
  } else {
    o.x = __fdiv_rn((float)p[0], 65535.0f);
    o.y = __fdiv_rn((float)p[1], 65535.0f);
    o.z = __fdiv_rn((float)p[2], 65535.0f);
  }
  o.w = 0.0f;
  reinterpret_cast<float4 *>(out)[idx] = o;
}

cudaError_t launch_gofloat_other(cudaStream_t s, int is16, const void *src, size_t owidth, size_t x, size_t y,
                                 size_t width, size_t height, const float2 *lut_rev, float *out) {
  size_t n = width * height;
  if (n == 0) return cudaSuccess;
  if (is16)
    k_gofloat_other<uint16_t><<<grid_for(n, 256), 256, 0, s>>>((const uint16_t *)src, owidth, x, y, width, height,
                                                               lut_rev, out);
  else
    k_gofloat_other<uint8_t><<<grid_for(n, 256), 256, 0, s>>>((const uint8_t *)src, owidth, x, y, width, height,
                                                              lut_rev, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ K2 demosaic::full (demosaic.rs:67-119)

// One CTA = a strip of 256 columns x kDemRows rows.  Every thread walks down its column with the 3x3 neighbourhood
// in registers (three new loads per pixel, consecutive lanes on consecutive addresses) and writes one float4 per pixel.
// Which taps feed which colour comes from a per-pattern-position mask table built once per CTA (demosaic.rs:77-90);
// out-of-frame taps are dropped from sum and count (:103-107); a colour without taps stays 0.0 (:110-114).
constexpr int kDemRows = 64;

__device__ __forceinline__ float dem_bin(uint32_t m, const float v[9]) {
  if (m == 0u) return 0.0f;
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; i++)
    if ((m >> i) & 1u) s = s + v[i];
  const int n = __popc(m);
  // s / 2^k == s * 2^-k for every float (an exact scaling, rounded once either way): Bayer frames never divide
  if ((n & (n - 1)) == 0) return s * __int_as_float((127 - (31 - __clz(n))) << 23);
  return __fdiv_rn(s, (float)n);
}

__global__ void __launch_bounds__(256)
k_demosaic_full(const __grid_constant__ CfaDev cfa, const float *__restrict__ in, int w, int h,
                float *__restrict__ out) {
  __shared__ uint2 taps[144];
  const int pw = cfa.width, ph = cfa.height;
  for (int pos = threadIdx.x; pos < pw * ph; pos += blockDim.x) {
    const int pr = pos / pw, pc = pos - pr * pw;
    const int pix = cfa.pat[pr * 48 + pc];
    uint32_t m[4] = {0u, 0u, 0u, 0u};
    int i = 0;
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++, i++) {
        const int oc = cfa.pat[((pr + 48 + dy) % 48) * 48 + (pc + 48 + dx) % 48];
        if ((oc != pix || (dx == 0 && dy == 0)) && oc < 4) m[oc] |= 1u << i;
      }
    taps[pos] = make_uint2(m[0] | (m[1] << 16), m[2] | (m[3] << 16));
  }
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= w) return;
  // RGB Bayer (2 x 2, green on one diagonal, red and blue on the other): the colours of positions (row & 1, col & 1) in
  // bits 2 * (2 * (row & 1) + (col & 1)), plus bit 8 so that the word is non-zero; 0 for every other pattern
  uint32_t bayer_phase = 0u;
  if (pw == 2 && ph == 2) {
    const int c0 = cfa.pat[0], c1 = cfa.pat[1], c2 = cfa.pat[48], c3 = cfa.pat[49];
    const bool ok = (c1 == 1 && c2 == 1 && ((c0 == 0 && c3 == 2) || (c0 == 2 && c3 == 0))) ||
                    (c0 == 1 && c3 == 1 && ((c1 == 0 && c2 == 2) || (c1 == 2 && c2 == 0)));
    if (ok) bayer_phase = 0x100u | (uint32_t)c0 | ((uint32_t)c1 << 2) | ((uint32_t)c2 << 4) | ((uint32_t)c3 << 6);
  }
  const int row0 = blockIdx.y * kDemRows, row1 = min(h, row0 + kDemRows);
  const int pc = col % pw;
  int pr = row0 % ph;
  const bool has_l = col > 0, has_r = col < w - 1;
  uint32_t colmask = 0x1ffu;
  if (!has_l) colmask &= ~0x049u;
  if (!has_r) colmask &= ~0x124u;
  auto load3 = [&](int r, float &a, float &b, float &c) {
    a = b = c = 0.0f;
    if (r >= 0 && r < h) {
      const float *p = in + (size_t)r * w + col;
      b = __ldg(p);
      if (has_l) a = __ldg(p - 1);
      if (has_r) c = __ldg(p + 1);
    }
  };
  float v[9];
  load3(row0 - 1, v[0], v[1], v[2]);
  load3(row0, v[3], v[4], v[5]);
  for (int row = row0; row < row1; row++) {
    load3(row + 1, v[6], v[7], v[8]);
    uint32_t valid = colmask;
    if (row == 0) valid &= ~0x007u;
    if (row == h - 1) valid &= ~0x1c0u;
    const uint2 mm = taps[pr * pw + pc];
    pr = pr + 1 == ph ? 0 : pr + 1;
    float4 o;
    if (bayer_phase != 0u && valid == 0x1ffu) {
      // RGB Bayer, all nine taps inside the frame: the four means a site can need, then selects (both kinds of site run
      // the same instructions).  Sums start at +0.0 and add in raster order like the reference's (demosaic.rs:96-114;
      // 0.0 + x also turns a -0.0 sample into the +0.0 the reference's sum holds); / 4, / 2 and / 1 are exact scalings.
      const float mg = ((((0.0f + v[1]) + v[3]) + v[5]) + v[7]) * 0.25f;
      const float md = ((((0.0f + v[0]) + v[2]) + v[6]) + v[8]) * 0.25f;
      const float mh = ((0.0f + v[3]) + v[5]) * 0.5f, mv = ((0.0f + v[1]) + v[7]) * 0.5f, own = 0.0f + v[4];
      const int sh = 2 * (2 * (row & 1) + (col & 1));
      const int c = (bayer_phase >> sh) & 3, ch = (bayer_phase >> (sh ^ 2)) & 3;  // this site's colour, its row neighbours'
      const bool gsite = c == 1;
      const int first = gsite ? ch : c;  // the colour (0 or 2) that receives a0
      const float a0 = gsite ? mh : own, a1 = gsite ? mv : md;
      o = make_float4(first == 0 ? a0 : a1, gsite ? own : mg, first == 0 ? a1 : a0, 0.0f);
    } else {
      o.x = dem_bin(mm.x & 0xffffu & valid, v);
      o.y = dem_bin((mm.x >> 16) & valid, v);
      o.z = dem_bin(mm.y & 0xffffu & valid, v);
      o.w = dem_bin((mm.y >> 16) & valid, v);
    }
    reinterpret_cast<float4 *>(out)[(size_t)row * w + col] = o;
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = v[i + 3];
  }
}

cudaError_t launch_demosaic_full(cudaStream_t s, const CfaDev &cfa, const float *in, size_t w, size_t h, float *out) {
  if (w == 0 || h == 0) return cudaSuccess;
  if (cfa.width <= 0 || cfa.height <= 0 || cfa.width * cfa.height > 144) return cudaErrorInvalidValue;
  if ((h + kDemRows - 1) / kDemRows > 65535) return cudaErrorInvalidValue;  // 4 M rows
  dim3 grid((unsigned)((w + 255) / 256), (unsigned)((h + kDemRows - 1) / kDemRows));
  k_demosaic_full<<<grid, 256, 0, s>>>(cfa, in, (int)w, (int)h, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ K3 transform_buffer<T> (scaling.rs:51-130)

struct XformDev {
  float tl0, tl1, skip_x_x, skip_x_y, skip_y_x, skip_y_y;
  size_t width, height, nwidth, nheight;
  int components;
};

template <typename T> __device__ __forceinline__ float as_f32(T v) { return (float)v; }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
// Rust `as u8` / `as u16`: saturating, NaN -> 0 (cvt.rzi.u32.f32 has the same semantics)
template <> __device__ __forceinline__ uint8_t from_f32<uint8_t>(float v) { return (uint8_t)min(__float2uint_rz(v), 255u); }
template <> __device__ __forceinline__ uint16_t from_f32<uint16_t>(float v) { return (uint16_t)min(__float2uint_rz(v), 65535u); }

__device__ __forceinline__ size_t f2usize(float f) { return (size_t)__float2ull_rz(f); }  // saturating, NaN -> 0

template <typename T, bool CFA>
__global__ void k_transform_buffer(const XformDev g, const __grid_constant__ CfaDev cfa, const T *__restrict__ src,
                                   T *__restrict__ out) {
  __shared__ uint8_t cfa_pat[CFA ? 48 * 48 : 1];
  if (CFA) {
    for (int i = threadIdx.x; i < 48 * 48; i += blockDim.x) cfa_pat[i] = cfa.pat[i];
    __syncthreads();
  }
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nwidth * g.nheight) return;
  size_t row = idx / g.nwidth, col = idx - row * g.nwidth;
  const float frow = (float)row, frow1 = (float)(row + 1), fcol = (float)col, fcol1 = (float)(col + 1);
  // per-row terms (scaling.rs:77-82)
  float rfrom_x = g.tl0 + g.skip_y_x * frow;
  float rto_x = g.tl0 + g.skip_y_x * frow1;
  float rfrom_y = g.tl1 + g.skip_y_y * frow;
  float rto_y = g.tl1 + g.skip_y_y * frow1;
  float rcenter_x = g.tl0 + (g.skip_y_x * frow) + __fdiv_rn(g.skip_y_x, 2.0f) - 0.5f;
  float rcenter_y = g.tl1 + (g.skip_y_y * frow) + __fdiv_rn(g.skip_y_y, 2.0f) - 0.5f;
  // per-pixel window (scaling.rs:84-89)
  size_t from_x = min(g.width - 1, f2usize(floorf(rfrom_x + (g.skip_x_x * fcol))));
  size_t to_x = min(g.width - 1, f2usize(floorf(rto_x + (g.skip_x_x * fcol1))));
  size_t from_y = min(g.height - 1, f2usize(floorf(rfrom_y + (g.skip_x_y * fcol))));
  size_t to_y = min(g.height - 1, f2usize(floorf(rto_y + (g.skip_x_y * fcol1))));
  float center_x = rcenter_x + (g.skip_x_x * fcol) + __fdiv_rn(g.skip_x_x, 2.0f);
  float center_y = rcenter_y + (g.skip_x_y * fcol) + __fdiv_rn(g.skip_x_y, 2.0f);

  float sums[4] = {0.f, 0.f, 0.f, 0.f}, counts[4] = {0.f, 0.f, 0.f, 0.f};
  const int comps = g.components;
  for (size_t y = from_y; y <= to_y; y++) {
    float delta_y = __fdiv_rn((float)y - center_y, g.skip_y_y);
    float dy2 = delta_y * delta_y;
    for (size_t x = from_x; x <= to_x; x++) {
      float delta_x = __fdiv_rn((float)x - center_x, g.skip_x_x);
      float factor = 1.0f - (delta_x * delta_x) - dy2;
      factor = factor < 0.0f ? 0.0f : factor;
      if (CFA) {
        int c = cfa_pat[(y % 48) * 48 + (x % 48)];
        float v = as_f32(src[y * g.width + x]) * factor;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (c == k) { sums[k] += v; counts[k] += factor; }
      } else {
        const T *p = src + (y * g.width + x) * comps;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (k < comps) { sums[k] += as_f32(p[k]) * factor; counts[k] += factor; }
      }
    }
  }
  T *o = out + idx * comps;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (k < comps) o[k] = counts[k] > 0.0f ? from_f32<T>(__fdiv_rn(sums[k], counts[k])) : from_f32<T>(0.0f);
}

static XformDev make_xform(const XformGeom &g) {
  XformDev d;
  d.tl0 = (float)g.tl[0];
  d.tl1 = (float)g.tl[1];
  // scaling.rs:69-72 — f32 host arithmetic (this file's host code is built with -ffp-contract=off)
  d.skip_x_x = ((float)g.tr[0] - (float)g.tl[0]) / (float)(g.nwidth - 1);
  d.skip_x_y = ((float)g.tr[1] - (float)g.tl[1]) / (float)(g.nwidth - 1);
  d.skip_y_x = ((float)g.bl[0] - (float)g.tl[0]) / (float)(g.nheight - 1);
  d.skip_y_y = ((float)g.bl[1] - (float)g.tl[1]) / (float)(g.nheight - 1);
  d.width = g.width; d.height = g.height; d.nwidth = g.nwidth; d.nheight = g.nheight;
  d.components = (int)g.components;
  return d;
}

static const CfaDev kNoCfa = {};

cudaError_t launch_transform_f32(cudaStream_t s, const XformGeom &g, const CfaDev *cfa, const float *src, float *out) {
  size_t n = g.nwidth * g.nheight;
  if (n == 0) return cudaSuccess;
  XformDev d = make_xform(g);
  if (cfa)
    k_transform_buffer<float, true><<<grid_for(n, 128), 128, 0, s>>>(d, *cfa, src, out);
  else
    k_transform_buffer<float, false><<<grid_for(n, 128), 128, 0, s>>>(d, kNoCfa, src, out);
  return cudaGetLastError();
}
cudaError_t launch_transform_u8(cudaStream_t s, const XformGeom &g, const uint8_t *src, uint8_t *out) {
  size_t n = g.nwidth * g.nheight;
  if (n == 0) return cudaSuccess;
  k_transform_buffer<uint8_t, false><<<grid_for(n, 128), 128, 0, s>>>(make_xform(g), kNoCfa, src, out);
  return cudaGetLastError();
}
cudaError_t launch_transform_u16(cudaStream_t s, const XformGeom &g, const uint16_t *src, uint16_t *out) {
  size_t n = g.nwidth * g.nheight;
  if (n == 0) return cudaSuccess;
  k_transform_buffer<uint16_t, false><<<grid_for(n, 128), 128, 0, s>>>(make_xform(g), kNoCfa, src, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ K4 to_lab (colorspaces.rs:89-112)

// Persistent CTAs (three per SM: 64 KB of shared memory each for the table) walk the buffer grid-stride; the table
// look-ups hit shared memory instead of gathering from L1/L2.
constexpr int kLutThreads = 512;
__device__ __forceinline__ void load_lut_smem(float2 *dst, const float2 *__restrict__ src) {
  const uint4 *a = reinterpret_cast<const uint4 *>(src);
  uint4 *d = reinterpret_cast<uint4 *>(dst);
  for (int i = threadIdx.x; i < kLutEntries / 2; i += blockDim.x) d[i] = __ldg(a + i);
  __syncthreads();
}
struct LutSmemPtr {
  const float2 *t;
  __device__ __forceinline__ float2 at(int key) const { return t[key]; }
};

// Lab transfer table in shared memory plus the context's table of the host libm's cbrtf over (1, 1.5] — the ratios that
// clipped highlights produce — so that only values beyond it (and -0.0 / NaN) take the double-precision restatement
struct LutSmemCbrt {
  const float2 *t;
  const float *cbrt_tab;
  __device__ __forceinline__ float2 at(int key) const { return t[key]; }
};
__device__ __forceinline__ float lab_f(const LutSmemCbrt &lut, float v) {
  if (v < 0.0f || v > 1.0f) {   // the same split as lab_f<Lut> (ipb_device.cuh)
    const uint32_t u = __float_as_uint(v), first = 0x3f800001u, size = 1u << 22;
    if (lut.cbrt_tab && u - first < size) return __ldg(lut.cbrt_tab + (u - first));
    return lab_f_slow(v);
  }
  return lut_lerp(lut, v);
}

// SplineFunc::interpolate (curves.rs:126-157) for finite, strictly increasing knots and finite coefficients (checked on
// the host: spline_counting_ok): the number of knots <= val names the piece — 0: below the first knot (its y), n: at or
// above the last (its y), else segment idx - 1, where a value on the knot itself gives y + 0 * c = y like the reference's
// early return.  NaN fails every comparison of the reference's binary search, which then returns y[(nseg - 1) / 2]
// (idx == 0 for NaN, so the NaN test comes first).
__device__ __forceinline__ float spline_sorted(const float (*spl)[8], const SplineDev &sp, float val) {
  int idx = 0;
  for (int j = 0; j < sp.n; j++) idx += val >= sp.x[j] ? 1 : 0;
  const float4 c = *reinterpret_cast<const float4 *>(spl[idx]);
  const float c3 = spl[idx][4];
  const float diff = val - c.x;
  const float r = c.y + c.z * diff + c.w * diff * diff + c3 * diff * diff * diff;
  // the end pieces return their y as it is (curves.rs:128-135): 0 * diff would be NaN for an infinite value, which these
  // kernels, unlike the fused ones, can meet
  if (val != val) return sp.y_nan;
  if (idx == 0 || idx == sp.n) return c.y;
  return r;
}

// CURVE != 0: OpBaseCurve (curves.rs:38-47: the spline on L, a and b untouched) applied to the pixel before it is
// stored — the two ops as one pass when nobody asked for the buffer between them (Pipeline::run without a cache).
// 1: the reference's binary search (any knots); 2: sorted knots, piece table in shared memory.
template <int CURVE>
__global__ void __launch_bounds__(kLutThreads)
k_tolab(const __grid_constant__ ColorParams P, const float2 *__restrict__ lut_lab, const float *__restrict__ cbrt_tab,
        const float *__restrict__ in, size_t npix, float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char lut_smem[];
  float2 *tab = reinterpret_cast<float2 *>(lut_smem);
  float (*spl)[8] = reinterpret_cast<float (*)[8]>(lut_smem + kLutEntries * sizeof(float2));
  if (CURVE == 2) {
    for (int i = threadIdx.x; i <= P.sp.n; i += kLutThreads) {
      float e[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      if (i == 0) e[1] = P.sp.y_first;
      else if (i >= P.sp.n) e[1] = P.sp.y_last;
      else { e[0] = P.sp.x[i - 1]; e[1] = P.sp.y[i - 1]; e[2] = P.sp.c1[i - 1]; e[3] = P.sp.c2[i - 1]; e[4] = P.sp.c3[i - 1]; }
      for (int k = 0; k < 5; k++) spl[i][k] = e[k];
    }
  }
  load_lut_smem(tab, lut_lab);
  const LutSmemCbrt lab{tab, cbrt_tab};
  // the next pixel is requested before this one is worked on (a pixel is ~230 instructions behind one 16-byte load)
  const size_t step = (size_t)gridDim.x * kLutThreads;
  size_t idx = (size_t)blockIdx.x * kLutThreads + threadIdx.x;
  float4 nxt = idx < npix ? __ldg(reinterpret_cast<const float4 *>(in) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (; idx < npix; idx += step) {
    const float4 px = nxt;
    if (idx + step < npix) nxt = __ldg(reinterpret_cast<const float4 *>(in) + idx + step);
    float l, a, b;
    camera_to_lab<false>(P, lab, px.x, px.y, px.z, px.w, l, a, b);
    if (CURVE == 1) l = spline_eval(P.sp, l);
    if (CURVE == 2) l = spline_sorted(spl, P.sp, l);
    out[idx * 3 + 0] = l;
    out[idx * 3 + 1] = a;
    out[idx * 3 + 2] = b;
  }
}
static int lut_grid(size_t work_items) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t blocks = (work_items + kLutThreads - 1) / kLutThreads;
  const size_t cap = (size_t)sms * 3;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}
template <int CURVE>
static cudaError_t launch_tolab_kind(cudaStream_t s, const ColorParams &P, const float2 *lut_lab, const float *cbrt_tab,
                                     const float *in, size_t npix, float *out) {
  const size_t smem = kLutEntries * sizeof(float2) + (kMaxSplinePts + 2) * 8 * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k_tolab<CURVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_tolab<CURVE><<<lut_grid(npix), kLutThreads, smem, s>>>(P, lut_lab, cbrt_tab, in, npix, out);
  return cudaGetLastError();
}
cudaError_t launch_tolab(cudaStream_t s, const ColorParams &P, const float2 *lut_lab, const float *cbrt_tab, int curve,
                         const float *in, size_t npix, float *out) {
  if (npix == 0) return cudaSuccess;
  if (curve == 2) return launch_tolab_kind<2>(s, P, lut_lab, cbrt_tab, in, npix, out);
  if (curve == 1) return launch_tolab_kind<1>(s, P, lut_lab, cbrt_tab, in, npix, out);
  return launch_tolab_kind<0>(s, P, lut_lab, cbrt_tab, in, npix, out);
}

// ------------------------------------------------------------------ K5 basecurve (curves.rs:33-49)

__global__ void k_basecurve(const __grid_constant__ SplineDev sp, const float *__restrict__ in, size_t npix,
                            float *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix) return;
  out[idx * 3 + 0] = spline_eval(sp, in[idx * 3 + 0]);
  out[idx * 3 + 1] = in[idx * 3 + 1];
  out[idx * 3 + 2] = in[idx * 3 + 2];
}
cudaError_t launch_basecurve(cudaStream_t s, const SplineDev &sp, const float *in, size_t npix, float *out) {
  if (npix == 0) return cudaSuccess;
  k_basecurve<<<grid_for(npix, 256), 256, 0, s>>>(sp, in, npix, out);
  return cudaGetLastError();
}

__global__ void k_spline_eval(const __grid_constant__ SplineDev sp, const float *__restrict__ in, size_t n,
                              float *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) out[idx] = spline_eval(sp, in[idx]);
}
cudaError_t launch_spline_eval(cudaStream_t s, const SplineDev &sp, const float *in, size_t n, float *out) {
  if (n == 0) return cudaSuccess;
  k_spline_eval<<<grid_for(n, 256), 256, 0, s>>>(sp, in, n, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ K6 from_lab (colorspaces.rs:127-137)

__global__ void k_fromlab(const __grid_constant__ ColorParams P, const float *__restrict__ in, size_t npix,
                          float *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix) return;
  float r, g, b;
  lab_to_rgb<false>(P, in[idx * 3 + 0], in[idx * 3 + 1], in[idx * 3 + 2], r, g, b);
  out[idx * 3 + 0] = r;
  out[idx * 3 + 1] = g;
  out[idx * 3 + 2] = b;
}
// from_lab and OpGamma (gamma.rs:21) as one pass when nobody asked for the buffer between them
__global__ void __launch_bounds__(kLutThreads)
k_fromlab_gamma(const __grid_constant__ ColorParams P, const float2 *__restrict__ lut_gamma, const float *__restrict__ in,
                size_t npix, float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char lut_smem[];
  float2 *tab = reinterpret_cast<float2 *>(lut_smem);
  load_lut_smem(tab, lut_gamma);
  const LutSmemPtr gam{tab};
  const size_t step = (size_t)gridDim.x * kLutThreads;
  size_t idx = (size_t)blockIdx.x * kLutThreads + threadIdx.x;
  float n0 = 0.f, n1 = 0.f, n2 = 0.f;
  if (idx < npix) { n0 = in[idx * 3 + 0]; n1 = in[idx * 3 + 1]; n2 = in[idx * 3 + 2]; }
  for (; idx < npix; idx += step) {
    const float l = n0, a = n1, bb = n2;
    if (idx + step < npix) { n0 = in[(idx + step) * 3 + 0]; n1 = in[(idx + step) * 3 + 1]; n2 = in[(idx + step) * 3 + 2]; }
    float r, g, b;
    lab_to_rgb<false>(P, l, a, bb, r, g, b);
    out[idx * 3 + 0] = gamma_elem(gam, r);
    out[idx * 3 + 1] = gamma_elem(gam, g);
    out[idx * 3 + 2] = gamma_elem(gam, b);
  }
}
cudaError_t launch_fromlab(cudaStream_t s, const ColorParams &P, const float2 *lut_gamma, const float *in, size_t npix,
                           float *out) {
  if (npix == 0) return cudaSuccess;
  if (lut_gamma) {
    const size_t smem = kLutEntries * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(k_fromlab_gamma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_fromlab_gamma<<<lut_grid(npix), kLutThreads, smem, s>>>(P, lut_gamma, in, npix, out);
  } else {
    k_fromlab<<<grid_for(npix, 256), 256, 0, s>>>(P, in, npix, out);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ K7 gamma (gamma.rs:16-26)

__global__ void __launch_bounds__(kLutThreads)
k_gamma(const float2 *__restrict__ lut_gamma, const float *__restrict__ in, size_t nelem, float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char lut_smem[];
  float2 *tab = reinterpret_cast<float2 *>(lut_smem);
  load_lut_smem(tab, lut_gamma);
  const LutSmemPtr gam{tab};
  const size_t nvec = nelem / 4;
  for (size_t i = (size_t)blockIdx.x * kLutThreads + threadIdx.x; i < nvec; i += (size_t)gridDim.x * kLutThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(in) + i);
    reinterpret_cast<float4 *>(out)[i] =
        make_float4(gamma_elem(gam, v.x), gamma_elem(gam, v.y), gamma_elem(gam, v.z), gamma_elem(gam, v.w));
  }
  for (size_t i = nvec * 4 + (size_t)blockIdx.x * kLutThreads + threadIdx.x; i < nelem; i += (size_t)gridDim.x * kLutThreads)
    out[i] = gamma_elem(gam, in[i]);
}
cudaError_t launch_gamma(cudaStream_t s, const float2 *lut_gamma, const float *in, size_t nelem, float *out) {
  if (nelem == 0) return cudaSuccess;
  const size_t smem = kLutEntries * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(k_gamma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_gamma<<<lut_grid(nelem / 4 + 1), kLutThreads, smem, s>>>(lut_gamma, in, nelem, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ K8 pack (pipeline.rs:408-414,455-461)

__global__ void k_pack8(const float *__restrict__ in, size_t nelem, uint8_t *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nelem) return;
  out[idx] = (uint8_t)output8bit(in[idx]);
}
__global__ void k_pack16(const float *__restrict__ in, size_t nelem, uint16_t *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nelem) return;
  out[idx] = (uint16_t)output16bit(in[idx]);
}
cudaError_t launch_pack8(cudaStream_t s, const float *in, size_t nelem, uint8_t *out) {
  if (nelem == 0) return cudaSuccess;
  k_pack8<<<grid_for(nelem, 256), 256, 0, s>>>(in, nelem, out);
  return cudaGetLastError();
}
cudaError_t launch_pack16(cudaStream_t s, const float *in, size_t nelem, uint16_t *out) {
  if (nelem == 0) return cudaSuccess;
  k_pack16<<<grid_for(nelem, 256), 256, 0, s>>>(in, nelem, out);
  return cudaGetLastError();
}

// image 0.24 DynamicImage::to_rgb8 of a 16-bit raster / to_rgb16 of an 8-bit raster (pipeline.rs:384,431; the
// crate is not in the reference tree: (v + 128) / 257 and v * 257 are its published conversions; unpinned)
__global__ void k_rgb16_to_8(const uint16_t *__restrict__ in, size_t n, uint8_t *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) out[idx] = (uint8_t)(((uint32_t)in[idx] + 128u) / 257u);
}
__global__ void k_rgb8_to_16(const uint8_t *__restrict__ in, size_t n, uint16_t *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) out[idx] = (uint16_t)((uint32_t)in[idx] * 257u);
}
cudaError_t launch_rgb16_to_8(cudaStream_t s, const uint16_t *in, size_t n, uint8_t *out) {
  if (n == 0) return cudaSuccess;
  k_rgb16_to_8<<<grid_for(n, 256), 256, 0, s>>>(in, n, out);
  return cudaGetLastError();
}
cudaError_t launch_rgb8_to_16(cudaStream_t s, const uint8_t *in, size_t n, uint16_t *out) {
  if (n == 0) return cudaSuccess;
  k_rgb8_to_16<<<grid_for(n, 256), 256, 0, s>>>(in, n, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ K9 rotate_buffer (transform.rs:87-144)
// Pure gather copy of 3-channel pixels; bit-exact by construction.  When transposing, a 32x32 pixel tile
// goes through shared memory so that both the global reads and the global writes are row-contiguous.

__global__ void k_rotate(const float *__restrict__ in, long sw, long sh, int transpose, long base_offset, long x_step,
                         long y_step, long ow, long oh, float *__restrict__ out) {
  __shared__ float tile[32][32 * 3 + 1];
  long tx0 = (long)blockIdx.x * 32, ty0 = (long)blockIdx.y * 32;  // output tile origin
  if (!transpose) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      long row = ty0 + r;
      if (row >= oh) break;
      long line_offset = base_offset + y_step * row;
      for (int e = threadIdx.x; e < 96; e += blockDim.x) {
        long col = tx0 + e / 3;
        int c = e % 3;
        if (col < ow) out[(row * ow + col) * 3 + c] = in[line_offset + x_step * col + c];
      }
    }
    return;
  }
  // transpose: output (row, col) reads source pixel at offset base + y_step*row + x_step*col where x_step is
  // a whole source row: source rows run along output columns.  Load the tile by source rows.
  for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {  // output col == one source row
    long col = tx0 + cc;
    if (col >= ow) break;
    for (int e = threadIdx.x; e < 96; e += blockDim.x) {  // output row == position along the source row
      int rr = e / 3, c = e % 3;
      long row = ty0 + rr;
      // y_step is +-3 here, so for flipped reads walk the source row backwards
      if (row < oh) tile[cc][rr * 3 + c] = in[base_offset + y_step * row + x_step * col + c];
    }
  }
  __syncthreads();
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    long row = ty0 + rr;
    if (row >= oh) break;
    for (int e = threadIdx.x; e < 96; e += blockDim.x) {
      int cc = e / 3, c = e % 3;
      long col = tx0 + cc;
      if (col < ow) out[(row * ow + col) * 3 + c] = tile[cc][rr * 3 + c];
    }
  }
}

cudaError_t launch_rotate(cudaStream_t s, const float *in, size_t w, size_t h, int transpose, int flip_x, int flip_y,
                          float *out) {
  if (w == 0 || h == 0) return cudaSuccess;
  long width = (long)w, height = (long)h;
  long base_offset = 0, x_step = 3, y_step = width * 3;
  if (flip_x) { x_step = -x_step; base_offset += (width - 1) * 3; }
  if (flip_y) { y_step = -y_step; base_offset += width * (height - 1) * 3; }
  long ow = width, oh = height;
  if (transpose) {
    ow = height; oh = width;
    long t = x_step; x_step = y_step; y_step = t;
  }
  dim3 block(32, 8), grid((unsigned)((ow + 31) / 32), (unsigned)((oh + 31) / 32));
  k_rotate<<<grid, block, 0, s>>>(in, width, height, transpose, base_offset, x_step, y_step, ow, oh, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ synthetic CFA frames

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_synth(uint64_t seed, size_t width, size_t row0, size_t rows, uint16_t *__restrict__ out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= width * rows) return;
  uint64_t i = (uint64_t)(row0 * width + idx);
  out[idx] = (uint16_t)(splitmix64(seed ^ i) & 16383u);
}
cudaError_t launch_synth(cudaStream_t s, uint64_t seed, size_t width, size_t row0, size_t rows, uint16_t *out) {
  size_t n = width * rows;
  if (n == 0) return cudaSuccess;
  k_synth<<<grid_for(n, 256), 256, 0, s>>>(seed, width, row0, rows, out);
  return cudaGetLastError();
}

}  // namespace ipb
