// ipb_fused.cu — the fused raw -> sRGB kernels (the roofline kernels of the hot path).
//
//   k_fused_full    gofloat (gofloat.rs:122-130) + demosaic::full (demosaic.rs:67-119) + to_lab + basecurve +
//                   from_lab + gamma (+ output8bit/output16bit pack) for a full-resolution CFA frame:
//                   u16 in (2 B/px), u8x3 / u16x3 / f32x3 out, nothing else touches HBM.
//   k_fused_scaled  the same chain with scaled_demosaic (scaling.rs:132-145) in place of full() — the branch
//                   OpDemosaic::run takes for scale >= 2 (Bayer) / 3 (X-Trans), demosaic.rs:47-50.
//
// k_fused_full is a persistent kernel: one 1024-thread CTA per SM walks 128x32-pixel tiles round-robin, one
// __syncthreads per tile.
//   * the raw u16 tile (+1 px halo, 144x34 box) of the NEXT tile is fetched by TMA (cp.async.bulk.tensor.2d,
//     completion on an mbarrier) while the current tile is computed;
//   * a conversion phase applies gofloat once per sensor pixel (smem u16 -> smem f32, double buffered; warps take
//     256-sample chunks from a counter, so the warps that finish their pixels first convert the next tile);
//   * the compute phase gives every thread four consecutive pixels: 3x6 window from shared memory, demosaic
//     (fixed-count means for RGB Bayer interiors, per-position tap masks for every other pattern and the borders),
//     then the colour chain on two pixel pairs in packed f32x2 arithmetic;
//   * both 8192-entry tables stay in shared memory for the whole launch.  For 8-bit output the gamma table is
//     replaced by a threshold table that yields output8bit(gamma(v)) directly (see build in ipb_host.cu);
//   * XYZ ratios outside [0,1] (the reference's analytic branch): cube roots of ratios in (1, 1.5] come from a table
//     of the host libm's cbrtf over every float of that range, negative ratios are evaluated inline, and whatever is
//     left (beyond 1.5, -0.0, NaN) is compacted into a per-warp queue and evaluated by the double-precision
//     restatement of glibc's cbrtf.
// k_fused_scaled gives every thread one output pixel: its window of raw samples is read from global memory
// (L1/L2 serve the overlap), weights per column / row are computed once, colours accumulate in the reference's tap
// order; RGB Bayer frames use a loop without colour look-ups.
// Arithmetic is the per-pixel code of ipb_device.cuh, compiled -fmad=false: results are bit-identical to the
// reference's f32 arithmetic (tests/test_gpu_fused*.py compare against the CPU oracle bit for bit).
#include <cuda.h>  // CUtensorMap (types only; cuTensorMapEncodeTiled is resolved at run time, no libcuda link)

#include <algorithm>
#include <cmath>
#include <cstring>

#include "ipb_internal.h"
#include "ipb_scaled.cuh"

namespace ipb {

namespace {

constexpr int kTW = 128;                  // tile width in pixels: 32 four-pixel tasks per tile row = one warp per row
constexpr int kTH = 32;                   // tile height (TMA boxes are at most 256 elements per dimension)
constexpr int kNT = 1024;                 // threads per CTA: one four-pixel task per thread per tile
constexpr int kWarps = kNT / 32;
constexpr int kTileStride = kTW + 16;     // elements per tile row: frame col tx0-8 .. tx0+kTW+7 (col tx0 at index 8); TMA needs
                                          // the box's first column on a 16-byte boundary (probed: tools/microbench/tma_probe.cu)
constexpr int kTileRows = kTH + 2;        // + one halo row above and below
constexpr int kTileElems = kTileRows * kTileStride;
constexpr int kStageElems = (kTileElems * 2 + 127) / 128 * 64;  // raw stage padded to a 128-byte multiple
constexpr int kMaxPatPos = 144;           // largest CFA period (12x12)
constexpr int kQueueCap = 12 * 32;        // out-of-table queue: 4 pixels x 3 ratios per lane
constexpr uint32_t kFull = 0xffffffffu;

struct FullParams {
  const uint16_t *raw;
  long long raw_pitch;
  int src_row0, src_rows;       // un-cropped full-frame rows present in raw
  int crop_x, crop_y;
  int width, height;            // cropped frame
  int out_row0, out_row1;
  void *out;
  float black, range, range_rc;
  int exact_rc;
  const float2 *lut_lab, *lut_out;
  const float *cbrt_tab;        // cbrtf of every float in (1.0, 1.5], indexed by bit pattern - 0x3f800001
  int tiles_x, tiles_y;
  int pw, ph;                   // CFA period
  uint32_t rcp_pw, rcp_ph;      // floor(2^32 / period) + 1 for the multiply-high remainder, 0: use the % operator
  int use_tma;
  int gamma8;                   // lut_out is the 8-bit threshold table
  uint32_t g8_bias;             // 0x4B000000 << 3 (mod 2^32), as a run-time value: ptxas would otherwise split the
                                // constant off the shift-add that forms a threshold-table address (two instructions)
};

struct Smem {
  float2 lut_lab[kLutEntries];
  float2 lut_out[kLutEntries];  // gamma {v, dv}; 8-bit output: {threshold, base} (Gamma8Entry)
  float tile[2][kTileElems];    // gofloat'ed tile, double buffered: tile t+1 is converted while tile t is computed
  alignas(128) uint16_t raw[kStageElems];  // TMA destination: raw u16 box of the next tile
  float queue[kWarps][kQueueCap];
  float spl[kMaxSplinePts + 2][8];  // [0] below the first knot, [1 + i] segment i, [n] at/above the last: x, y, c1, c2, c3
  uint2 taps[kMaxPatPos];       // per pattern position: 9-bit tap masks of colours 0..3, 16 bits each
  alignas(8) unsigned long long mbar;
  alignas(8) unsigned long long mbar_lut;      // completion of the two bulk copies that bring the tables in
  int conv_ctr[2];              // dynamic chunk counters of the conversion phase (alternating per tile)
};
static_assert(sizeof(Smem) <= 232448, "Smem exceeds the 227 KB a CTA can opt in to");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- TMA + mbarrier (PTX)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}

// demosaic.rs:77-90 for every position of the CFA period: which of the nine 3x3 taps feed which colour bin.
// Taps of the centre's own colour other than the centre itself go to the discarded fifth bin.
template <class S>
__device__ __forceinline__ void build_taps(S &sm, const CfaDev &cfa, int pw, int ph) {
  for (int pos = threadIdx.x; pos < pw * ph; pos += blockDim.x) {
    int pr = pos / pw, pc = pos - pr * pw;
    int pix = cfa.pat[pr * 48 + pc];
    uint32_t m[4] = {0u, 0u, 0u, 0u};
    int i = 0;
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++, i++) {
        int oc = cfa.pat[((pr + 48 + dy) % 48) * 48 + (pc + 48 + dx) % 48];
        if ((oc != pix || (dx == 0 && dy == 0)) && oc < 4) m[oc] |= 1u << i;
      }
    sm.taps[pos] = make_uint2(m[0] | (m[1] << 16), m[2] | (m[3] << 16));
  }
}

// One colour bin of demosaic::full for one pixel: taps summed in the reference's order (-1,-1)..(1,1),
// skipped taps contribute +0.0 (x + 0.0 == x for every partial sum, which is never -0.0), divide by the count.
__device__ __forceinline__ float bin_mean(uint32_t m, const float v[9]) {
  if (m == 0u) return 0.0f;
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; i++) s = s + (((m >> i) & 1u) ? v[i] : 0.0f);
  return __fdiv_rn(s, (float)__popc(m));
}

// {1/n rounded to nearest, n} for tap counts n = 0..9 (entry 0 divides the empty sum by 1: +0.0)
__constant__ float2 kTapRcp[10] = {{0.0f, 1.0f}, {1.0f, 1.0f}, {0.5f, 2.0f}, {1.0f / 3.0f, 3.0f}, {0.25f, 4.0f},
                                   {1.0f / 5.0f, 5.0f}, {1.0f / 6.0f, 6.0f}, {1.0f / 7.0f, 7.0f}, {0.125f, 8.0f},
                                   {1.0f / 9.0f, 9.0f}};

// The same bin for a pixel whose nine taps are all inside the frame, without the IEEE division: the sum over the
// selected taps in the reference's order (predicated adds), then s / n through the 3-instruction reciprocal form,
// which equals IEEE division for every divisor 1..9 (tools/verify_constdiv.c; sums of level-mapped samples are zero
// or normal numbers).
__device__ __forceinline__ float bin_mean_rc(uint32_t m, const float v[9]) {
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; i++)
    if ((m >> i) & 1u) s = s + v[i];
  const float2 e = kTapRcp[__popc(m)];
  return div_rc(s, e.y, e.x);
}

template <int OUT>
__device__ __forceinline__ void store_px4(void *out, size_t pix_index, int n, const float r[4], const float g[4],
                                          const float b[4]) {
  if (OUT == kOutF32) {
    float *o = reinterpret_cast<float *>(out) + pix_index * 3;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      float4 *o4 = reinterpret_cast<float4 *>(o);
      o4[0] = make_float4(r[0], g[0], b[0], r[1]);
      o4[1] = make_float4(g[1], b[1], r[2], g[2]);
      o4[2] = make_float4(b[2], r[3], g[3], b[3]);
    } else {
      for (int j = 0; j < n; j++) { o[j * 3] = r[j]; o[j * 3 + 1] = g[j]; o[j * 3 + 2] = b[j]; }
    }
  } else if (OUT == kOutU8) {
    uint32_t q[12];
#pragma unroll
    for (int j = 0; j < 4; j++) { q[j * 3] = output8bit(r[j]); q[j * 3 + 1] = output8bit(g[j]); q[j * 3 + 2] = output8bit(b[j]); }
    uint8_t *o = reinterpret_cast<uint8_t *>(out) + pix_index * 3;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
      uint32_t *o4 = reinterpret_cast<uint32_t *>(o);
      o4[0] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
      o4[1] = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
      o4[2] = q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24);
    } else {
      for (int j = 0; j < n * 3; j++) o[j] = (uint8_t)q[j];
    }
  } else {
    uint32_t q[12];
#pragma unroll
    for (int j = 0; j < 4; j++) { q[j * 3] = output16bit(r[j]); q[j * 3 + 1] = output16bit(g[j]); q[j * 3 + 2] = output16bit(b[j]); }
    uint16_t *o = reinterpret_cast<uint16_t *>(out) + pix_index * 3;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
      uint2 *o2 = reinterpret_cast<uint2 *>(o);
      o2[0] = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
      o2[1] = make_uint2(q[4] | (q[5] << 16), q[6] | (q[7] << 16));
      o2[2] = make_uint2(q[8] | (q[9] << 16), q[10] | (q[11] << 16));
    } else {
      for (int j = 0; j < n * 3; j++) o[j] = (uint16_t)q[j];
    }
  }
}

// 8-bit output already quantised (q = 12 bytes in 12 registers)
__device__ __forceinline__ void store_px4_bytes(void *out, size_t pix_index, int n, const uint32_t q[12]) {
  uint8_t *o = reinterpret_cast<uint8_t *>(out) + pix_index * 3;
  if (n == 4 && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
    uint32_t *o4 = reinterpret_cast<uint32_t *>(o);
    o4[0] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
    o4[1] = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
    o4[2] = q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24);
  } else {
    for (int j = 0; j < n * 3; j++) o[j] = (uint8_t)q[j];
  }
}

// ---------------------------------------------------------------- colour chain on a pixel pair
// The pair (pixel j, pixel j+1) of a thread's four pixels travels through to_lab, basecurve, from_lab and gamma
// as the two halves of packed f32x2 registers (FMUL2 / FFMA2): same IEEE results as the scalar code of
// ipb_device.cuh with half the issue slots for the arithmetic.  Table look-ups, min/max and selects are per half.

// true when lookup() takes the table branch with a key the masked lerp computes correctly: +0.0 <= v <= 1.0
__device__ __forceinline__ bool in_table(float v) { return __float_as_uint(v) <= 0x3f800000u; }

struct LerpIdx {
  F2 a;   // pos - trunc(pos)
  F2 tf;  // 2^23 + floor(pos): the key sits in the low mantissa bits
};
// index part of TransformLookup::lookup for both halves: pos = v*max; key = trunc(pos); a = pos - trunc(pos)
__device__ __forceinline__ LerpIdx lerp_index(const PkAdd &pk, F2 v) {
  LerpIdx r;
  const F2 pos = pk_mul(v, kLutMax);
  r.tf = pk.add_rm(pos, 8388608.0f);
  const F2 base = pk.add(r.tf, -8388608.0f);
  r.a = pk.sub(pos, base);
  return r;
}
// table part: v1 + a*(v2 - v1) from the {v1, v2 - v1} entry; the key is masked so that any bit pattern reads a valid
// entry (values outside the table get their result from lab_f_slow() instead)
__device__ __forceinline__ float lerp_fetch(uint32_t lut_base, float tf, float a) {
  const uint32_t off = (__float_as_uint(tf) << 3) & 0xfff8u;
  float2 e;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e.x), "=f"(e.y) : "r"(lut_base + off));
  return e.x + a * e.y;
}

// SplineFunc::interpolate (curves.rs:126-157) from the table in shared memory.  Entry = number of knots <= val: 0 is
// "at or below the first knot" and n "at or above the last" (curves.rs:128-135), both stored as constant cubics
// {y, 0, 0, 0} — y + 0*d + 0*d*d + 0*d*d*d == y for every finite d — and entry 1 + i is segment i.  On a knot the
// cubic of the segment that starts there gives y exactly, like the reference's early return.  The fused launch is
// gated on finite inputs and strictly increasing knots (ipb_host.cu fused_params_bounded), so NaN never gets here.
__device__ __forceinline__ float spline_eval_smem(const float (*spl)[8], const SplineDev &s, float val) {
  int idx;
  if (s.n == 3) {  // the raw default: one interior knot (curves.rs:14-20)
    idx = val >= s.x[1] ? (val >= s.x[2] ? 3 : 2) : (val >= s.x[0] ? 1 : 0);
  } else {
    idx = (val >= s.x[0] ? 1 : 0) + (val >= s.x[1] ? 1 : 0);
    for (int j = 2; j < s.n; j++) idx += val >= s.x[j] ? 1 : 0;
  }
  const float4 c = *reinterpret_cast<const float4 *>(spl[idx]);
  const float c3 = spl[idx][4];
  const float diff = val - c.x;
  return c.y + c.z * diff + c.w * diff * diff + c3 * diff * diff * diff;
}

// 8-bit output of one channel: output8bit(apply_srgb_gamma(clamp(v))) through the threshold table.  vc is clamped
// to [0,1], so tf = 2^23 + key with key <= 8191: its bit pattern is 0x4B000000 + key and (bits << 3) + bias, with
// bias = table base - (0x4B000000 << 3 mod 2^32), is the entry's address in one shift-add.
__device__ __forceinline__ uint32_t gamma8_fetch(uint32_t lut_bias, float tf, float vc) {
  float thr;
  uint32_t base;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(thr), "=r"(base) : "r"((__float_as_uint(tf) << 3) + lut_bias));
  return base + (vc >= thr ? 1u : 0u);
}
__device__ __forceinline__ F2 clamp01(F2 v) {
  return F2{fminf(fmaxf(v.x, 0.0f), 1.0f), fminf(fmaxf(v.y, 0.0f), 1.0f)};
}

// camera_to_lab up to the XYZ ratios (color_conversions.rs:42-55,156-160) for two pixels
__device__ __forceinline__ void xyz_ratios_pair(const ColorParams &P, const PkAdd &pk, const float r[2], const float g[2],
                                                const float b[2], const float e[2], F2 &xr, F2 &yr, F2 &zr) {
  // white balance, clip, 3x4 matrix with left-to-right sums
  F2 cr = pk_mul(F2{r[0], r[1]}, P.mul[0]);
  F2 cg = pk_mul(F2{g[0], g[1]}, P.mul[1]);
  F2 cb = pk_mul(F2{b[0], b[1]}, P.mul[2]);
  cr = F2{fminf(cr.x, 1.0f), fminf(cr.y, 1.0f)};
  cg = F2{fminf(cg.x, 1.0f), fminf(cg.y, 1.0f)};
  cb = F2{fminf(cb.x, 1.0f), fminf(cb.y, 1.0f)};
  F2 x = pk.add(pk.add(pk_mul(cr, P.cm[0]), pk_mul(cg, P.cm[1])), pk_mul(cb, P.cm[2]));
  F2 y = pk.add(pk.add(pk_mul(cr, P.cm[4]), pk_mul(cg, P.cm[5])), pk_mul(cb, P.cm[6]));
  F2 z = pk.add(pk.add(pk_mul(cr, P.cm[8]), pk_mul(cg, P.cm[9])), pk_mul(cb, P.cm[10]));
  if (P.use_e) {
    F2 ce = pk_mul(F2{e[0], e[1]}, P.mul[3]);
    ce = F2{fminf(ce.x, 1.0f), fminf(ce.y, 1.0f)};
    x = pk.add(x, pk_mul(ce, P.cm[3]));
    y = pk.add(y, pk_mul(ce, P.cm[7]));
    z = pk.add(z, pk_mul(ce, P.cm[11]));
  }
  xr = IPB_PK_DIVC(x, 0.95047f);
  yr = y;  // y / 1.0
  zr = IPB_PK_DIVC(z, 1.08883f);
}

// XYZ_LAB_TRANSFORM.lookup, table branch, of the four values of one channel (pixels 0..3 of the task).  Values
// outside the table read some valid entry (masked key) and are replaced afterwards (lab_outside_table).
__device__ __forceinline__ void lab_lookup4(const PkAdd &pk, uint32_t lab_base, F2 va, F2 vb, float f[4]) {
  const LerpIdx ia = lerp_index(pk, va), ib = lerp_index(pk, vb);
  f[0] = lerp_fetch(lab_base, ia.tf.x, ia.a.x);
  f[1] = lerp_fetch(lab_base, ia.tf.y, ia.a.y);
  f[2] = lerp_fetch(lab_base, ib.tf.x, ib.a.x);
  f[3] = lerp_fetch(lab_base, ib.tf.y, ib.a.y);
}

constexpr uint32_t kCbrtTabFirst = 0x3f800001u;  // bit pattern of the first tabulated value: the float after 1.0
constexpr uint32_t kCbrtTabSize = 1u << 22;      // ... up to and including 1.5 (0x3fc00000)

// The reference's analytic branch of XYZ_LAB_TRANSFORM.lookup (color_conversions.rs:102-104,120-124) for the task's
// twelve XYZ ratios v (x of pixels 0..3, y, z), entered when some lane of the warp holds a ratio outside [+0, 1]:
//   1 < v <= 1.5   v.cbrt(): read from a table of the host libm's cbrtf over every float of that range (16 MB, built
//                  once per context, L2-resident) — ratios a little above 1 are what clipped highlights produce;
//   v < 0          (k*v + 16) / 116, computed for every value in packed arithmetic and selected;
//   anything else  (v > 1.5, -0.0, NaN): compacted into the warp's queue and evaluated by lab_f_slow (glibc's cbrtf
//                  restated in double precision); practically never taken, the fused launch is gated on finite inputs.
__device__ __forceinline__ void lab_outside_table(const PkAdd &pk, const float *__restrict__ cbrt_tab, float *queue, int lane,
                                                  const F2 vp[6], float f[12]) {
  const float kk = 24389.0f / 27.0f;
  float neg[12];
#pragma unroll
  for (int h = 0; h < 6; h++) {
    const F2 n = IPB_PK_DIVC(pk.add(pk_mul(vp[h], kk), 16.0f), 116.0f);
    neg[2 * h] = n.x;
    neg[2 * h + 1] = n.y;
  }
  uint32_t wmin = 0xffffffffu, umax = 0u;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const float v = (i & 1) ? vp[i >> 1].y : vp[i >> 1].x;
    const uint32_t u = __float_as_uint(v);
    const uint32_t idx = u - kCbrtTabFirst;
    if (idx < kCbrtTabSize) f[i] = __ldg(cbrt_tab + idx);
    if (v < 0.0f) f[i] = neg[i];
    wmin = min(wmin, u - (kCbrtTabFirst + kCbrtTabSize));  // 0 .. 0x403fffff for 1.5 < v, +inf, +NaN and -0.0
    umax = max(umax, u);
  }
  const bool rest = wmin <= 0x80000000u - (kCbrtTabFirst + kCbrtTabSize) || umax > 0xff800000u;  // or a negative NaN
  if (__any_sync(kFull, rest)) {
    int cnt = 0;
    bool mine[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const uint32_t u = __float_as_uint((i & 1) ? vp[i >> 1].y : vp[i >> 1].x);
      mine[i] = u - (kCbrtTabFirst + kCbrtTabSize) <= 0x80000000u - (kCbrtTabFirst + kCbrtTabSize) || u > 0xff800000u;
      cnt += mine[i] ? 1 : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, d);
      incl += lane >= d ? t : 0;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    int q = incl - cnt;
#pragma unroll
    for (int i = 0; i < 12; i++)
      if (mine[i]) queue[q++] = (i & 1) ? vp[i >> 1].y : vp[i >> 1].x;
    __syncwarp();
    for (int i = lane; i < total; i += 32) queue[i] = lab_f_slow(queue[i]);
    __syncwarp();
    q = incl - cnt;
#pragma unroll
    for (int i = 0; i < 12; i++)
      if (mine[i]) f[i] = queue[q++];
    __syncwarp();
  }
}

// Lab from the transfer-function values, basecurve, from_lab (+ gamma) for two pixels
template <int OUT>
__device__ __forceinline__ void lab_to_output_pair(const ColorParams &P, const PkAdd &pk, const Smem &sm,
                                                   uint32_t out_base, uint32_t g8_bias, bool g8, F2 fx, F2 fy, F2 fz,
                                                   float orr[2],
                                                   float og[2], float ob[2], uint32_t q8[6]) {
  F2 l = pk.add(pk_mul(fy, 116.0f), -16.0f);
  F2 a = pk_mul(pk.sub(fx, fy), 500.0f);
  F2 bb = pk_mul(pk.sub(fy, fz), 200.0f);
  l = IPB_PK_DIVC(l, 100.0f);
  a = IPB_PK_DIVC(pk.add(a, 127.0f), 255.0f);
  bb = IPB_PK_DIVC(pk.add(bb, 127.0f), 255.0f);
  // ---- basecurve on L (curves.rs:45-47)
  if (P.sp.n > 0) {
    l.x = spline_eval_smem(sm.spl, P.sp, l.x);
    l.y = spline_eval_smem(sm.spl, P.sp, l.y);
  }
  // ---- lab_to_xyz + lab_to_rgb (color_conversions.rs:58-65,172-191)
  const float ee = 216.0f / 24389.0f;
  const float kk = 24389.0f / 27.0f;
  const F2 cl = pk_mul(l, 100.0f);
  const F2 ca = pk.add(pk_mul(a, 255.0f), -127.0f);
  const F2 cbb = pk.add(pk_mul(bb, 255.0f), -127.0f);
  const F2 ffy = IPB_PK_DIVC(pk.add(cl, 16.0f), 116.0f);
  const F2 ffx = pk.add(IPB_PK_DIVC(ca, 500.0f), ffy);
  const F2 ffz = pk.sub(ffy, IPB_PK_DIVC(cbb, 200.0f));
  const F2 fx3 = pk_mul(pk_mul(ffx, ffx), ffx);
  const F2 fy3 = pk_mul(pk_mul(ffy, ffy), ffy);
  const F2 fz3 = pk_mul(pk_mul(ffz, ffz), ffz);
  const F2 xlin = IPB_PK_DIVC(pk.add(pk_mul(ffx, 116.0f), -16.0f), kk);
  const F2 ylin = IPB_PK_DIVC(cl, kk);
  const F2 zlin = IPB_PK_DIVC(pk.add(pk_mul(ffz, 116.0f), -16.0f), kk);
  const float ke = kk * ee;
  const F2 xr2{fx3.x > ee ? fx3.x : xlin.x, fx3.y > ee ? fx3.y : xlin.y};
  const F2 yr2{cl.x > ke ? fy3.x : ylin.x, cl.y > ke ? fy3.y : ylin.y};
  const F2 zr2{fz3.x > ee ? fz3.x : zlin.x, fz3.y > ee ? fz3.y : zlin.y};
  const F2 X = pk_mul(xr2, 0.95047f);
  const F2 Y = yr2;  // * 1.0
  const F2 Z = pk_mul(zr2, 1.08883f);
  F2 rr = pk.add(pk.add(pk_mul(X, P.rgbm[0]), pk_mul(Y, P.rgbm[1])), pk_mul(Z, P.rgbm[2]));
  F2 gg = pk.add(pk.add(pk_mul(X, P.rgbm[3]), pk_mul(Y, P.rgbm[4])), pk_mul(Z, P.rgbm[5]));
  F2 bl = pk.add(pk.add(pk_mul(X, P.rgbm[6]), pk_mul(Y, P.rgbm[7])), pk_mul(Z, P.rgbm[8]));
  // ---- gamma (gamma.rs:21) and quantisation
  if (OUT == kOutU8 && g8) {
    const F2 vr = clamp01(rr), vg = clamp01(gg), vb = clamp01(bl);
    const F2 tr = pk.add_rm(pk_mul(vr, kLutMax), 8388608.0f);
    const F2 tg = pk.add_rm(pk_mul(vg, kLutMax), 8388608.0f);
    const F2 tb = pk.add_rm(pk_mul(vb, kLutMax), 8388608.0f);
    const uint32_t bias = out_base - g8_bias;
    q8[0] = gamma8_fetch(bias, tr.x, vr.x);
    q8[1] = gamma8_fetch(bias, tg.x, vg.x);
    q8[2] = gamma8_fetch(bias, tb.x, vb.x);
    q8[3] = gamma8_fetch(bias, tr.y, vr.y);
    q8[4] = gamma8_fetch(bias, tg.y, vg.y);
    q8[5] = gamma8_fetch(bias, tb.y, vb.y);
  } else {
    if (!P.linear) {
      const LerpIdx jr = lerp_index(pk, clamp01(rr)), jg = lerp_index(pk, clamp01(gg)), jb = lerp_index(pk, clamp01(bl));
      rr = F2{lerp_fetch(out_base, jr.tf.x, jr.a.x), lerp_fetch(out_base, jr.tf.y, jr.a.y)};
      gg = F2{lerp_fetch(out_base, jg.tf.x, jg.a.x), lerp_fetch(out_base, jg.tf.y, jg.a.y)};
      bl = F2{lerp_fetch(out_base, jb.tf.x, jb.a.x), lerp_fetch(out_base, jb.tf.y, jb.a.y)};
    }
    if (OUT == kOutU8) {
      q8[0] = output8bit(rr.x); q8[1] = output8bit(gg.x); q8[2] = output8bit(bl.x);
      q8[3] = output8bit(rr.y); q8[4] = output8bit(gg.y); q8[5] = output8bit(bl.y);
    }
    orr[0] = rr.x; orr[1] = rr.y; og[0] = gg.x; og[1] = gg.y; ob[0] = bl.x; ob[1] = bl.y;
  }
}

// one thread: arm the stage's barrier with the box size and start the bulk tensor copy of tile (txi, tyi)
__device__ __forceinline__ void issue_tile(const FullParams &p, const CUtensorMap *tmap, uint32_t raw_stage,
                                           uint32_t bar, int txi, int tyi) {
  const int x = txi * kTW - 8 + p.crop_x;  // multiple of 8 samples: the launcher checks crop_x % 8 == 0
  const int y = p.out_row0 + tyi * kTH - 1 + p.crop_y - p.src_row0;
  mbar_expect_tx(bar, kTileElems * (uint32_t)sizeof(uint16_t));
  tma_load_2d(raw_stage, tmap, x, y, bar);
}

// ---------------------------------------------------------------------------------------- full resolution

template <int OUT, bool BAYER>
__global__ void __launch_bounds__(kNT, 1)
k_fused_full(const __grid_constant__ FullParams p, const __grid_constant__ CfaDev cfa,
             const __grid_constant__ ColorParams P, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  const int ntiles = p.tiles_x * p.tiles_y;
  const uint32_t bar = smem_u32(&sm.mbar), raw_addr = smem_u32(sm.raw);

  int tyi = (int)blockIdx.x / p.tiles_x, txi = (int)blockIdx.x - tyi * p.tiles_x;
  const int step_y = (int)gridDim.x / p.tiles_x, step_x = (int)gridDim.x - step_y * p.tiles_x;
  auto next_tile = [&](int &tx, int &ty) {
    tx += step_x; ty += step_y;
    if (tx >= p.tiles_x) { tx -= p.tiles_x; ty++; }
  };

  // Both 64 KB tables come in by bulk asynchronous copy (the TMA engine, no tensor map), issued by one thread and
  // overlapped with the first tile's copy, the tap / spline tables and the first conversion; everybody waits for them
  // just before the first look-up.
  const uint32_t bar_lut = smem_u32(&sm.mbar_lut);
  if (tid == 0) {
    sm.conv_ctr[0] = 0;
    sm.conv_ctr[1] = 0;
    mbar_init(bar, 1);
    mbar_init(bar_lut, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (p.use_tma && (int)blockIdx.x < ntiles) issue_tile(p, &tmap, raw_addr, bar, txi, tyi);
    constexpr uint32_t kLutBytes = kLutEntries * (uint32_t)sizeof(float2);
    mbar_expect_tx(bar_lut, 2 * kLutBytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sm.lut_lab)), "l"(p.lut_lab), "r"(kLutBytes), "r"(bar_lut) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sm.lut_out)), "l"(p.lut_out), "r"(kLutBytes), "r"(bar_lut) : "memory");
  }
  build_taps(sm, cfa, p.pw, p.ph);
  for (int i = tid; i < kMaxSplinePts + 2; i += kNT) {
    float e[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (i == 0) e[1] = P.sp.y_first;
    else if (i >= P.sp.n) e[1] = P.sp.y_last;
    else { e[0] = P.sp.x[i - 1]; e[1] = P.sp.y[i - 1]; e[2] = P.sp.c1[i - 1]; e[3] = P.sp.c2[i - 1]; e[4] = P.sp.c3[i - 1]; }
#pragma unroll
    for (int k = 0; k < 5; k++) sm.spl[i][k] = e[k];
  }
  // shared-window addresses of the tables, made opaque so that they stay in registers: the compiler otherwise
  // re-derives them from the CTA id (three uniform-pipe instructions) in front of every group of look-ups
  uint32_t lab_base = smem_u32(sm.lut_lab), out_base = smem_u32(sm.lut_out);
  asm volatile("" : "+r"(lab_base), "+r"(out_base));
  const int lane = tid & 31, warp = tid >> 5;
  float *queue = sm.queue[warp];
  const bool g8 = OUT == kOutU8 && p.gamma8 != 0;

  // gofloat (gofloat.rs:122-130) once per sensor pixel, into tile buffer `buf`.  TMA path: the raw box is in shared
  // memory; warps pull 32-thread chunks (eight samples per thread) from a counter so that whichever warps finish
  // the previous tile's pixels first do the conversion (task times vary with the data: the out-of-table queue; a
  // static split was measured 4 % slower despite fewer instructions).  Plain path: bounds-checked global loads.
  // Also measured and dropped: replacing the CTA barrier by per-buffer mbarriers so that warps run up to a tile
  // apart (issue utilisation 0.73 -> 0.78, but the polling costs more than it gains: 88 vs 92 GP/s).
  auto convert_tile = [&](int buf, int ctr, int ttx0, int tty0) {
    float *tile = sm.tile[buf];
    if (p.use_tma) {
      constexpr int kChunks = (kTileElems / 8 + 31) / 32;
      for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = atomicAdd(&sm.conv_ctr[ctr], 1);
        chunk = __shfl_sync(kFull, chunk, 0);
        if (chunk >= kChunks) break;
        const int g8i = chunk * 32 + lane;
        if (g8i < kTileElems / 8) {
          const uint4 pkd = *reinterpret_cast<const uint4 *>(sm.raw + g8i * 8);
          const uint32_t w4[4] = {pkd.x, pkd.y, pkd.z, pkd.w};
          float v[8];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            v[2 * k] = golevel((float)(w4[k] & 0xffffu), p.black, p.range, p.range_rc, p.exact_rc);
            v[2 * k + 1] = golevel((float)(w4[k] >> 16), p.black, p.range, p.range_rc, p.exact_rc);
          }
          float4 *dst = reinterpret_cast<float4 *>(tile + g8i * 8);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    } else {
      for (int i = tid; i < kTileRows * (kTW + 2); i += kNT) {
        const int r = i / (kTW + 2), c = i - r * (kTW + 2);
        const int y = tty0 - 1 + r, x = ttx0 - 1 + c;
        float v = 0.0f;
        if (y >= 0 && y < p.height && x >= 0 && x < p.width) {
          const int sr = y + p.crop_y - p.src_row0;
          if (sr >= 0 && sr < p.src_rows) {
            const uint16_t rawv = __ldg(p.raw + (long long)sr * p.raw_pitch + p.crop_x + x);
            v = golevel((float)rawv, p.black, p.range, p.range_rc, p.exact_rc);
          }
        }
        tile[r * kTileStride + c + 7] = v;
      }
    }
  };

  // Bayer phase (cropped-frame coordinates): colour of the pixel at (row&1, col&1)
  const int c00 = cfa.pat[0], c10 = cfa.pat[48];

  // ---- prologue: tile 0 converted, tile 1 in flight
  __syncthreads();  // tables, counters and the barrier are ready
  if ((int)blockIdx.x < ntiles) {
    if (p.use_tma) mbar_wait(bar, 0);
    convert_tile(0, 0, txi * kTW, p.out_row0 + tyi * kTH);
  }
  __syncthreads();
  {
    int ntx = txi, nty = tyi;
    next_tile(ntx, nty);
    if (tid == 0 && p.use_tma && (int)(blockIdx.x + gridDim.x) < ntiles) issue_tile(p, &tmap, raw_addr, bar, ntx, nty);
  }

  mbar_wait(bar_lut, 0);  // the tables are in
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
    const int ty0 = p.out_row0 + tyi * kTH, tx0 = txi * kTW;
    next_tile(txi, tyi);  // (txi, tyi) now name tile t + gridDim.x
    const float *ctile = sm.tile[it & 1];

    // ---- four-pixel tasks: 64 per tile row, consecutive lanes on consecutive tasks
    for (int task = tid; task < (kTW / 4) * kTH; task += kNT) {
      const int r = task / (kTW / 4), q = task - r * (kTW / 4);
      const int y = ty0 + r, x0 = tx0 + 4 * q;
      const bool live = y < p.out_row1 && x0 < p.width;
      const int npx = live ? min(4, p.width - x0) : 0;
      const bool interior = y >= 1 && y <= p.height - 2 && x0 >= 1 && x0 + 4 <= p.width - 1;
      const bool fast = BAYER && __all_sync(kFull, !live || interior);

      // 3 x 6 window: rows y-1..y+1, cols x0-1..x0+4
      float w[3][6];
      const float *tp = ctile + r * kTileStride + 4 * q + 7;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float4 mid = *reinterpret_cast<const float4 *>(tp + k * kTileStride + 1);
        w[k][0] = tp[k * kTileStride];
        w[k][1] = mid.x; w[k][2] = mid.y; w[k][3] = mid.z; w[k][4] = mid.w;
        w[k][5] = tp[k * kTileStride + 5];  // (fetching the two halo columns from the neighbouring lanes by shuffle
      }                                     //  instead was measured slower: 252 vs 246 us per C2 frame)

      float cr[4], cg[4], cb[4], ce[4];
      if (fast) {
        // RGB Bayer interior: row colour pattern is (A, G, A, G ...) or (G, A, G, A ...), A in {R, B}
        const int cfirst = (y & 1) ? c10 : c00;      // colour of even columns on this row
        const bool green_first = cfirst == 1;
        const int a_col = green_first ? ((y & 1) ? cfa.pat[49] : cfa.pat[1]) : cfirst;  // the row's non-green colour
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const bool is_green = green_first ? ((j & 1) == 0) : ((j & 1) == 1);
          const float n = w[0][j + 1], s = w[2][j + 1], wv = w[1][j], e = w[1][j + 2], c = w[1][j + 1];
          float va, vg, vo;  // row colour, green, other colour
          if (is_green) {
            va = (wv + e) * 0.5f;
            vg = c;
            vo = (n + s) * 0.5f;
          } else {
            va = c;
            vg = (((n + wv) + e) + s) * 0.25f;
            vo = (((w[0][j] + w[0][j + 2]) + w[2][j]) + w[2][j + 2]) * 0.25f;
          }
          cr[j] = a_col == 0 ? va : vo;
          cg[j] = vg;
          cb[j] = a_col == 0 ? vo : va;
          ce[j] = 0.0f;
        }
      } else if (BAYER) {
        // RGB Bayer warp with pixels on the frame border: the same neighbourhoods, but a tap outside the frame is
        // dropped from sum and count (demosaic.rs:103-107).  Sums start at 0.0 and take the taps in the reference's
        // order with +0.0 for a dropped one (a no-op on a partial sum, which is never -0.0); the mean is s / count
        // through the exact reciprocal form for counts 1..4.  Pixels beyond the frame (partial tiles) compute on
        // whatever the tile holds and are not stored.
        const int cfirst = (y & 1) ? c10 : c00;
        const bool green_first = cfirst == 1;
        const int a_col = green_first ? ((y & 1) ? cfa.pat[49] : cfa.pat[1]) : cfirst;
        const bool has_n = y > 0, has_s = y < p.height - 1;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int x = x0 + j;
          const bool has_w = x > 0, has_e = x < p.width - 1;
          const bool is_green = green_first ? ((j & 1) == 0) : ((j & 1) == 1);
          const float c = w[1][j + 1];
          const float n = has_n ? w[0][j + 1] : 0.0f, s = has_s ? w[2][j + 1] : 0.0f;
          const float wv = has_w ? w[1][j] : 0.0f, e = has_e ? w[1][j + 2] : 0.0f;
          float va, vg, vo;
          if (is_green) {
            const float2 ra = kTapRcp[max((int)has_w + (int)has_e, 1)], ro = kTapRcp[max((int)has_n + (int)has_s, 1)];
            va = div_rc((0.0f + wv) + e, ra.y, ra.x);
            vg = c;
            vo = div_rc((0.0f + n) + s, ro.y, ro.x);
          } else {
            const float nw = (has_n && has_w) ? w[0][j] : 0.0f, ne = (has_n && has_e) ? w[0][j + 2] : 0.0f;
            const float sw = (has_s && has_w) ? w[2][j] : 0.0f, se = (has_s && has_e) ? w[2][j + 2] : 0.0f;
            const float2 rg = kTapRcp[max((int)has_n + (int)has_w + (int)has_e + (int)has_s, 1)];
            const float2 ro = kTapRcp[max(((int)has_n + (int)has_s) * ((int)has_w + (int)has_e), 1)];
            va = c;
            vg = div_rc((((0.0f + n) + wv) + e) + s, rg.y, rg.x);
            vo = div_rc((((0.0f + nw) + ne) + sw) + se, ro.y, ro.x);
          }
          cr[j] = a_col == 0 ? va : vo;
          cg[j] = vg;
          cb[j] = a_col == 0 ? vo : va;
          ce[j] = 0.0f;
        }
      } else {
        // generic CFA / frame border: per-position tap masks, out-of-frame taps dropped from sum and count.
        // With TMA staging the tile holds whatever lies outside the cropped frame; the masks never select it.
        // position inside the CFA period: n % d as n - mulhi(n, ceil(2^32 / d)) * d (exact while n * d < 2^32)
        const int pr = p.rcp_ph ? y - (int)__umulhi((uint32_t)y, p.rcp_ph) * p.ph : y % p.ph;
        int pc = p.rcp_pw ? x0 - (int)__umulhi((uint32_t)x0, p.rcp_pw) * p.pw : x0 % p.pw;
        const bool inside = __all_sync(kFull, !live || interior);  // no tap of the warp's pixels leaves the frame
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int x = x0 + j;
          const uint2 mm = sm.taps[pr * p.pw + pc];
          pc = (pc + 1 == p.pw) ? 0 : pc + 1;
          float v[9];
#pragma unroll
          for (int k = 0; k < 3; k++) { v[k * 3] = w[k][j]; v[k * 3 + 1] = w[k][j + 1]; v[k * 3 + 2] = w[k][j + 2]; }
          if (inside) {
            cr[j] = bin_mean_rc(mm.x & 0xffffu, v);
            cg[j] = bin_mean_rc(mm.x >> 16, v);
            cb[j] = bin_mean_rc(mm.y & 0xffffu, v);
            ce[j] = P.use_e ? bin_mean_rc(mm.y >> 16, v) : 0.0f;
          } else {
            uint32_t valid = 0x1ffu;
            if (y <= 0) valid &= ~0x007u;
            if (y >= p.height - 1) valid &= ~0x1c0u;
            if (x <= 0) valid &= ~0x049u;
            if (x >= p.width - 1) valid &= ~0x124u;
            cr[j] = bin_mean(mm.x & 0xffffu & valid, v);
            cg[j] = bin_mean((mm.x >> 16) & valid, v);
            cb[j] = bin_mean(mm.y & 0xffffu & valid, v);
            ce[j] = bin_mean((mm.y >> 16) & valid, v);
          }
        }
      }

      float orr[4], og[4], ob[4];
      uint32_t q8[12];
      // ---- colour chain: pixels (0,1) and (2,3) as packed pairs; one queue flush for the task's 12 table look-ups
      const PkAdd pk{P.one, P.mone};
      F2 xa, ya, za, xb, yb, zb;
      xyz_ratios_pair(P, pk, cr, cg, cb, ce, xa, ya, za);
      xyz_ratios_pair(P, pk, cr + 2, cg + 2, cb + 2, ce + 2, xb, yb, zb);
      float fl[12];  // transfer-function values in the order x0 x1 x2 x3 y0 .. z3
      lab_lookup4(pk, lab_base, xa, xb, fl);
      lab_lookup4(pk, lab_base, ya, yb, fl + 4);
      lab_lookup4(pk, lab_base, za, zb, fl + 8);
      {
        // any ratio of the warp outside [+0, 1]?  As unsigned bit patterns, negative values, values above 1.0 and NaN
        // all compare above 1.0f.
        const F2 vp[6] = {xa, xb, ya, yb, za, zb};
        uint32_t mx = 0u;
#pragma unroll
        for (int h = 0; h < 6; h++) mx = max(mx, max(__float_as_uint(vp[h].x), __float_as_uint(vp[h].y)));
        if (__any_sync(kFull, mx > 0x3f800000u)) lab_outside_table(pk, p.cbrt_tab, queue, lane, vp, fl);
      }
      const float *fxs = fl, *fys = fl + 4, *fzs = fl + 8;
      lab_to_output_pair<OUT>(P, pk, sm, out_base, p.g8_bias, g8, F2{fxs[0], fxs[1]}, F2{fys[0], fys[1]}, F2{fzs[0], fzs[1]}, orr, og,
                              ob, q8);
      lab_to_output_pair<OUT>(P, pk, sm, out_base, p.g8_bias, g8, F2{fxs[2], fxs[3]}, F2{fys[2], fys[3]}, F2{fzs[2], fzs[3]}, orr + 2,
                              og + 2, ob + 2, q8 + 6);
      if (live) {
        const size_t pix = (size_t)(y - p.out_row0) * p.width + x0;
        if (OUT == kOutU8) store_px4_bytes(p.out, pix, npx, q8);
        else store_px4<OUT>(p.out, pix, npx, orr, og, ob);
      }
    }

    // ---- convert the next tile (its raw box was requested one tile ago), then one barrier per tile
    const bool have_next = t + (int)gridDim.x < ntiles;
    if (have_next) {
      if (p.use_tma) mbar_wait(bar, (it + 1) & 1);
      convert_tile((it + 1) & 1, (it + 1) & 1, txi * kTW, p.out_row0 + tyi * kTH);
    }
    if (tid == 0) sm.conv_ctr[it & 1] = 0;  // idle during this tile (the conversion above counts on the other one)
    __syncthreads();  // tile t consumed, tile t+1 converted, raw box free again
    if (tid == 0) {
      if (p.use_tma && t + 2 * (int)gridDim.x < ntiles) {
        int ntx = txi, nty = tyi;
        next_tile(ntx, nty);
        issue_tile(p, &tmap, raw_addr, bar, ntx, nty);
      }
    }
  }
}

struct SmemScaled {
  float2 lut_lab[kLutEntries];
  float2 lut_gamma[kLutEntries];
  uint8_t pat[48 * kPatStride];
  alignas(8) unsigned long long mbar_lut;  // completion of the bulk copies that bring the two tables in
};

constexpr int kNTScaled = 1024;

// k_fused_scaled: one output pixel per thread (scaling.rs:76-127 with the CFA binning of :109-112), then the colour
// chain.  The reference's per-tap arithmetic is kept expression by expression; what is shared between taps is
// computed once: delta_x and 1 - delta_x^2 per window column (the reference recomputes them for every row),
// delta_y^2 per window row.  A colour's weighted sum and weight sum travel as one packed f32x2 accumulator:
// (v, 1) * (f, f) = (v*f, f) is one FMUL2 and the two additions one predicated packed add, each half rounded exactly
// like the reference's scalar `sums[c] += v*f; counts[c] += f` — and in the same tap order.  The window loop is
// unrolled for the warp's widest window (5 columns at 4x; 6 and 8 as fall-backs), without per-lane predication when
// every lane has that width.
template <int OUT>
__global__ void __launch_bounds__(kNTScaled, 1)
k_fused_scaled(const __grid_constant__ ScaledParams p, const __grid_constant__ CfaDev cfa,
               const __grid_constant__ ColorParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemScaled &sm = *reinterpret_cast<SmemScaled *>(smem_raw);
  // the tables come in by bulk asynchronous copy while the first windows are accumulated (they are first needed by
  // the colour chain)
  const uint32_t bar_lut = smem_u32(&sm.mbar_lut);
  if (threadIdx.x == 0) {
    mbar_init(bar_lut, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    constexpr uint32_t kLutBytes = kLutEntries * (uint32_t)sizeof(float2);
    mbar_expect_tx(bar_lut, 2 * kLutBytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sm.lut_lab)), "l"(p.lut_lab), "r"(kLutBytes), "r"(bar_lut) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sm.lut_gamma)), "l"(p.lut_gamma), "r"(kLutBytes), "r"(bar_lut) : "memory");
  }
  bool luts_in = false;
  bool has_e = false;
  for (int i = threadIdx.x; i < 48 * kPatStride; i += blockDim.x) {
    const int r = i / kPatStride, c = i - r * kPatStride;
    sm.pat[i] = cfa.pat[r * 48 + (c % 48)];
    has_e |= sm.pat[i] >= 3;
  }
  const bool four = __syncthreads_or(has_e) != 0;  // a fourth colour somewhere in the pattern
  const LutShared lab{smem_u32(sm.lut_lab)}, gam{smem_u32(sm.lut_gamma)};

  const long long npix = (long long)(p.out_row1 - p.out_row0) * p.nwidth;
  const long long npix_pad = (npix + 31) / 32 * 32;  // whole warps stay in the loop (warp-wide votes below)
  for (long long idx = (long long)blockIdx.x * kNTScaled + threadIdx.x; idx < npix_pad;
       idx += (long long)gridDim.x * kNTScaled) {
    const bool live = idx < npix;
    const long long id = live ? idx : npix - 1;
    const int row = p.out_row0 + (int)(id / p.nwidth), col = (int)(id % p.nwidth);
    const ScaledWindow win = scaled_window(p, row, col);
    const int from_x = win.from_x, to_x = win.to_x, from_y = win.from_y, to_y = win.to_y;
    const float center_x = win.center_x, center_y = win.center_y;
    const int nx = to_x - from_x + 1;

    F2 acc[4] = {F2{0.f, 0.f}, F2{0.f, 0.f}, F2{0.f, 0.f}, F2{0.f, 0.f}};  // {sums[c], counts[c]}
    const int nx_max = __reduce_max_sync(kFull, nx), nx_min = __reduce_min_sync(kFull, nx);
    if (nx_max <= kMaxCols && p.exact_rc && p.bayer) {
      if (nx_max == 5 && nx_min == 5) {
        if (p.skip_rc_exact) window_taps_bayer<5, true, true>(p, cfa, from_x, nx, from_y, to_y, center_x, center_y, acc);
        else window_taps_bayer<5, true, false>(p, cfa, from_x, nx, from_y, to_y, center_x, center_y, acc);
      } else if (nx_max <= 6) {
        window_taps_bayer<6, false, false>(p, cfa, from_x, nx, from_y, to_y, center_x, center_y, acc);
      } else {
        window_taps_bayer<kMaxCols, false, false>(p, cfa, from_x, nx, from_y, to_y, center_x, center_y, acc);
      }
    } else if (nx_max <= kMaxCols && p.exact_rc && !four) {
      if (nx_max == 5 && nx_min == 5)
        window_taps<5, true, 3>(p, sm.pat, P.one, from_x, nx, from_y, to_y, center_x, center_y, acc);
      else if (nx_max <= 6)
        window_taps<6, false, 3>(p, sm.pat, P.one, from_x, nx, from_y, to_y, center_x, center_y, acc);
      else
        window_taps<kMaxCols, false, 3>(p, sm.pat, P.one, from_x, nx, from_y, to_y, center_x, center_y, acc);
    } else if (nx_max <= kMaxCols && p.exact_rc) {
      window_taps<kMaxCols, false, 4>(p, sm.pat, P.one, from_x, nx, from_y, to_y, center_x, center_y, acc);
    } else {
      // wide windows (scale >= 7) or a level mapping that needs IEEE division: the plain loop
      for (int y = from_y; y <= to_y; y++) {
        const float delta_y = __fdiv_rn((float)y - center_y, p.skip_y);
        const float dy2 = delta_y * delta_y;
        const uint16_t *rowp = p.raw + (long long)(y + p.crop_y - p.src_row0) * p.raw_pitch + p.crop_x;
        const uint8_t *prow = sm.pat + (y % 48) * kPatStride;
        for (int x = from_x; x <= to_x; x++) {
          const float delta_x = __fdiv_rn((float)x - center_x, p.skip_x);
          float factor = 1.0f - (delta_x * delta_x) - dy2;
          factor = factor < 0.0f ? 0.0f : factor;
          const int c = prow[x % 48];
          const float v = golevel((float)__ldg(rowp + x), p.black, p.range, p.range_rc, p.exact_rc) * factor;
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (c == j) { acc[j].x += v; acc[j].y += factor; }
        }
      }
    }
    float px[4];
#pragma unroll
    for (int k = 0; k < 4; k++) px[k] = acc[k].y > 0.0f ? __fdiv_rn(acc[k].x, acc[k].y) : 0.0f;
    float r[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (!luts_in) {
      mbar_wait(bar_lut, 0);
      luts_in = true;
    }
    color_chain<true>(P, lab, gam, px[0], px[1], px[2], px[3], r[0], g[0], b[0]);
    if (live) store_px4<OUT>(p.out, (size_t)(idx), 1, r, g, b);
  }
}

// ---------------------------------------------------------------------------------------- gamma8 checks

// every non-negative f32 bit pattern up to 1.0, plus negative / >1 / NaN inputs via the sign and top bits
__global__ void k_gamma8_selftest(const float2 *__restrict__ lut_gamma, const float2 *__restrict__ lut_gamma8,
                                  unsigned long long *mismatches) {
  const LutGlobal gam{lut_gamma};
  unsigned long long bad = 0;
  const uint32_t nthreads = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t b = t0; b <= 0xffffffffull; b += nthreads) {
    const uint32_t bits = (uint32_t)b;
    // [0, 1.0] exhaustively; elsewhere (negative, > 1, inf, NaN) every 4099th pattern
    if (bits > 0x3f800000u && (bits % 4099u) != 0u) continue;
    const float v = __uint_as_float(bits);
    const uint32_t want = output8bit(gamma_elem(gam, v));
    const float vc = fminf(fmaxf(v, 0.0f), 1.0f);
    const float pos = vc * kLutMax;
    const float tf = __fadd_rd(pos, 8388608.0f);
    const float2 e = __ldg(lut_gamma8 + (__float_as_uint(tf) & 0x1fffu));
    const uint32_t got = __float_as_uint(e.y) + (vc >= e.x ? 1u : 0u);
    bad += got != want;
  }
  if (bad) atomicAdd(mismatches, bad);
}

__global__ void k_gamma8_pack(const float2 *__restrict__ lut_gamma8, const float *__restrict__ in, size_t n,
                              uint8_t *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float vc = fminf(fmaxf(in[i], 0.0f), 1.0f);
  const float tf = __fadd_rd(vc * kLutMax, 8388608.0f);
  const float2 e = __ldg(lut_gamma8 + (__float_as_uint(tf) & 0x1fffu));
  out[i] = (uint8_t)(__float_as_uint(e.y) + (vc >= e.x ? 1u : 0u));
}

thread_local const char *g_fused_err = "";

template <class K>
cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table: libipb200.so does not link libcuda,
// so it still loads (and exports its symbols) on a box without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// 2-D u16 tensor map over the source rows present in `raw` (un-cropped sensor width), box = one staged tile.
bool make_raw_tmap(CUtensorMap *map, const uint16_t *raw, size_t pitch_elems, size_t width_elems, size_t rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(raw) & 15) || ((pitch_elems * sizeof(uint16_t)) & 15) || rows == 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)width_elems, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)(pitch_elems * sizeof(uint16_t))};
  const cuuint32_t box[2] = {(cuuint32_t)kTileStride, (cuuint32_t)kTileRows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t *>(raw), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int OUT>
cudaError_t launch_full_kind(cudaStream_t s, const FullParams &p, const CfaDev &cfa, const ColorParams &P,
                             const CUtensorMap &tmap, int grid, bool bayer) {
  const size_t smem = sizeof(Smem);
  cudaError_t e;
  if (bayer) {
    if ((e = set_smem(k_fused_full<OUT, true>, smem)) != cudaSuccess) return e;
    k_fused_full<OUT, true><<<grid, kNT, smem, s>>>(p, cfa, P, tmap);
  } else {
    if ((e = set_smem(k_fused_full<OUT, false>, smem)) != cudaSuccess) return e;
    k_fused_full<OUT, false><<<grid, kNT, smem, s>>>(p, cfa, P, tmap);
  }
  return cudaGetLastError();
}

}  // namespace

const char *fused_last_error() { return g_fused_err; }

cudaError_t launch_gamma8_selftest(cudaStream_t s, const float2 *lut_gamma, const float2 *lut_gamma8,
                                   unsigned long long *mismatches) {
  k_gamma8_selftest<<<148 * 8, 256, 0, s>>>(lut_gamma, lut_gamma8, mismatches);
  return cudaGetLastError();
}
cudaError_t launch_gamma8_pack(cudaStream_t s, const float2 *lut_gamma8, const float *in, size_t n, uint8_t *out) {
  if (n == 0) return cudaSuccess;
  k_gamma8_pack<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(lut_gamma8, in, n, out);
  return cudaGetLastError();
}

static bool is_rgb_bayer(const CfaDev &cfa) {
  if (cfa.width != 2 || cfa.height != 2) return false;
  const int a = cfa.pat[0], b = cfa.pat[1], c = cfa.pat[48], d = cfa.pat[49];
  // green (colour 1) on one diagonal, red and blue on the other
  if (b == 1 && c == 1) return (a == 0 && d == 2) || (a == 2 && d == 0);
  if (a == 1 && d == 1) return (b == 0 && c == 2) || (b == 2 && c == 0);
  return false;
}

cudaError_t launch_fused_full(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                              int sm_count) {
  if (a.out_row1 <= a.out_row0 || a.width == 0) return cudaSuccess;
  if (cfa.width <= 0 || cfa.height <= 0 || cfa.width * cfa.height > kMaxPatPos) {
    g_fused_err = "fused_full: unsupported CFA period";
    return cudaErrorInvalidValue;
  }
  FullParams p;
  p.raw = a.raw; p.raw_pitch = (long long)a.raw_pitch;
  p.src_row0 = (int)a.src_row0; p.src_rows = (int)a.src_rows;
  p.crop_x = (int)a.crop_x; p.crop_y = (int)a.crop_y;
  p.width = (int)a.width; p.height = (int)a.height;
  p.out_row0 = (int)a.out_row0; p.out_row1 = (int)a.out_row1;
  p.out = a.out;
  p.black = a.black; p.range = a.range; p.range_rc = a.range_rc; p.exact_rc = a.exact_rc;
  p.lut_lab = a.lut_lab;
  p.cbrt_tab = a.cbrt_tab;
  p.gamma8 = (a.out_kind == kOutU8 && !P.linear && a.lut_gamma8) ? 1 : 0;
  p.lut_out = p.gamma8 ? a.lut_gamma8 : a.lut_gamma;
  p.g8_bias = 0x58000000u;
  p.tiles_x = (p.width + kTW - 1) / kTW;
  p.tiles_y = (p.out_row1 - p.out_row0 + kTH - 1) / kTH;
  p.pw = cfa.width; p.ph = cfa.height;
  const bool small = a.width < (1u << 26) && a.height < (1u << 26);  // n * period < 2^32 for every coordinate
  p.rcp_pw = small && p.pw > 1 ? (uint32_t)(0x100000000ull / (uint64_t)p.pw) + 1u : 0u;
  p.rcp_ph = small && p.ph > 1 ? (uint32_t)(0x100000000ull / (uint64_t)p.ph) + 1u : 0u;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  p.use_tma = (a.use_tma && (a.crop_x % 8) == 0 && make_raw_tmap(&tmap, a.raw, a.raw_pitch, a.raw_pitch, a.src_rows)) ? 1 : 0;
  const bool bayer = is_rgb_bayer(cfa);
  const int ntiles = p.tiles_x * p.tiles_y;
  const int grid = ntiles < sm_count ? ntiles : sm_count;
  switch (a.out_kind) {
    case kOutF32: return launch_full_kind<kOutF32>(s, p, cfa, P, tmap, grid, bayer);
    case kOutU8: return launch_full_kind<kOutU8>(s, p, cfa, P, tmap, grid, bayer);
    default: return launch_full_kind<kOutU16>(s, p, cfa, P, tmap, grid, bayer);
  }
}

// Does the three-instruction reciprocal form of (coordinate - centre) / skip give the IEEE quotient for every tap of every
// window of this frame?  The kernel's coordinate arithmetic (k_fused_scaled, scaling.rs:77-89,94-98) is repeated here in
// the same f32 expressions, column by column and row by row.  A few ten thousand divisions per distinct frame geometry.
static bool scaled_skip_rc_exact(int width, int height, int nwidth, int nheight, float skip_x, float skip_y) {
  const float rcx = 1.0f / skip_x, rcy = 1.0f / skip_y;
  if (!std::isnormal(skip_x) || !std::isnormal(skip_y) || !std::isnormal(rcx) || !std::isnormal(rcy)) return false;
  auto sat = [](float f) { return f >= 2147483648.0f ? 0x7fffffff : (f > 0.0f ? (int)f : 0); };
  auto same = [](float num, float d, float rc) {
    const float q = num * rc, r = fmaf(-q, d, num), q2 = fmaf(r, rc, q), ref = num / d;
    return memcmp(&q2, &ref, 4) == 0;
  };
  for (int col = 0; col < nwidth; col++) {
    const float fcol = (float)col, fcol1 = (float)(col + 1);
    const float rcenter_x = 0.0f + (0.0f * 0.0f) + (0.0f / 2.0f) - 0.5f;
    const int from_x = std::min(width - 1, sat(floorf(0.0f + (skip_x * fcol))));
    const int to_x = std::min(width - 1, sat(floorf(0.0f + (skip_x * fcol1))));
    const float center_x = rcenter_x + (skip_x * fcol) + (skip_x / 2.0f);
    for (int x = from_x; x <= to_x; x++)
      if (!same((float)x - center_x, skip_x, rcx)) return false;
  }
  for (int row = 0; row < nheight; row++) {
    const float frow = (float)row, frow1 = (float)(row + 1);
    const float rcenter_y = 0.0f + (skip_y * frow) + (skip_y / 2.0f) - 0.5f;
    const int from_y = std::min(height - 1, sat(floorf((0.0f + skip_y * frow) + 0.0f)));
    const int to_y = std::min(height - 1, sat(floorf((0.0f + skip_y * frow1) + 0.0f)));
    const float center_y = rcenter_y + 0.0f + (0.0f / 2.0f);
    for (int y = from_y; y <= to_y; y++)
      if (!same((float)y - center_y, skip_y, rcy)) return false;
  }
  return true;
}

bool scaled_skip_division_exact(size_t width, size_t height, size_t nwidth, size_t nheight) {
  if (nwidth < 2 || nheight < 2 || width < 2 || height < 2 || width >= (1u << 30) || height >= (1u << 30)) return false;
  const float skip_x = ((float)((long)width - 1) - 0.0f) / (float)(nwidth - 1);
  const float skip_y = ((float)((long)height - 1) - 0.0f) / (float)(nheight - 1);
  return scaled_skip_rc_exact((int)width, (int)height, (int)nwidth, (int)nheight, skip_x, skip_y);
}

void fill_scaled_params(const FusedArgs &a, const CfaDev &cfa, ScaledParams *out) {
  ScaledParams p;
  p.raw = a.raw; p.raw_pitch = (long long)a.raw_pitch;
  p.src_row0 = (int)a.src_row0; p.src_rows = (int)a.src_rows;
  p.crop_x = (int)a.crop_x; p.crop_y = (int)a.crop_y;
  p.width = (int)a.width; p.height = (int)a.height;
  p.nwidth = (int)a.out_width; p.nheight = (int)a.out_height;
  p.out_row0 = (int)a.out_row0; p.out_row1 = (int)a.out_row1;
  p.out = a.out;
  p.black = a.black; p.range = a.range; p.range_rc = a.range_rc; p.exact_rc = a.exact_rc;
  p.lut_lab = a.lut_lab; p.lut_gamma = a.lut_gamma;
  // scaling.rs:46,69-72: corners (0,0), (width-1,0), (0,height-1)
  p.skip_x = ((float)((long)a.width - 1) - 0.0f) / (float)(a.out_width - 1);
  p.skip_y = ((float)((long)a.height - 1) - 0.0f) / (float)(a.out_height - 1);
  p.bayer = is_rgb_bayer(cfa) ? 1 : 0;
  p.skip_x_rc = 1.0f / p.skip_x;
  p.skip_y_rc = 1.0f / p.skip_y;
  {
    // the check depends on the frame geometry only: remembered for the last one (frames of a batch share it)
    struct Memo { int w, h, nw, nh, ok; };
    thread_local Memo memo = {0, 0, 0, 0, 0};
    if (memo.w != p.width || memo.h != p.height || memo.nw != p.nwidth || memo.nh != p.nheight) {
      memo = {p.width, p.height, p.nwidth, p.nheight,
              scaled_skip_rc_exact(p.width, p.height, p.nwidth, p.nheight, p.skip_x, p.skip_y) ? 1 : 0};
    }
    p.skip_rc_exact = memo.ok;
  }
  {
    const bool integral = a.black >= 0.0f && a.black < 4194304.0f && a.black == floorf(a.black);
    p.sub_a = integral ? -(8388608.0f + a.black) : -8388608.0f;
    p.sub_b = integral ? 0.0f : -a.black;
  }
  *out = p;
}

cudaError_t launch_fused_scaled(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                                int sm_count) {
  if (a.out_row1 <= a.out_row0 || a.out_width == 0) return cudaSuccess;
  ScaledParams p;
  fill_scaled_params(a, cfa, &p);
  const long long npix = (long long)(p.out_row1 - p.out_row0) * p.nwidth;
  long long blocks = (npix + kNTScaled - 1) / kNTScaled;
  const int grid = (int)(blocks < sm_count ? blocks : sm_count);
  const size_t smem = sizeof(SmemScaled);
  cudaError_t e;
  switch (a.out_kind) {
    case kOutF32:
      if ((e = set_smem(k_fused_scaled<kOutF32>, smem)) != cudaSuccess) return e;
      k_fused_scaled<kOutF32><<<grid, kNTScaled, smem, s>>>(p, cfa, P);
      break;
    case kOutU8:
      if ((e = set_smem(k_fused_scaled<kOutU8>, smem)) != cudaSuccess) return e;
      k_fused_scaled<kOutU8><<<grid, kNTScaled, smem, s>>>(p, cfa, P);
      break;
    default:
      if ((e = set_smem(k_fused_scaled<kOutU16>, smem)) != cudaSuccess) return e;
      k_fused_scaled<kOutU16><<<grid, kNTScaled, smem, s>>>(p, cfa, P);
      break;
  }
  return cudaGetLastError();
}

}  // namespace ipb
