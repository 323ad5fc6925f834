// ipb_device.cuh — per-pixel device arithmetic of the raw->sRGB hot path.
//
// Exactness discipline (DESIGN.md "Numerics"): this translation unit set is compiled with
// -fmad=false, so `a*b + c` is two IEEE roundings exactly like the reference (Rust never contracts
// to FMA).  fmaf() is used only inside div_rc(), a 3-instruction division by a constant that is
// proven bit-identical to IEEE division by exhaustion (tools/verify_constdiv.c).  All constants are
// f32-evaluated the way the reference evaluates them (color_conversions.rs:121-122,181-182).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipb {

constexpr int kLutEntries = 8192;     // float2 {table[i], table[i+1]-table[i]}, i = 0..8191
constexpr float kLutMax = 8191.0f;    // TransformLookup::max, color_conversions.rs:89
constexpr int kMaxSplinePts = 34;     // IPB_MAX_CURVE_POINTS + 2 auto-added end points

// SplineFunc coefficients (curves.rs:59-64), built on the host by the same f32 code path.
struct SplineDev {
  int n;     // number of points (0 => basecurve is a pass-through, curves.rs:34-36)
  int nseg;  // c3s.len()
  float x_first, y_first, x_last, y_last;  // curves.rs:128-135 end clamps
  float y_nan;                             // what the reference's binary search returns for NaN: y[(nseg-1)/2]
  float x[kMaxSplinePts], y[kMaxSplinePts], c1[kMaxSplinePts], c2[kMaxSplinePts], c3[kMaxSplinePts];
};

// Everything the colour chain needs, derived on the host from OpToLab/OpBaseCurve/settings.
struct ColorParams {
  float mul[4];    // normalize_wbs(wb_coeffs) or 1s for monochrome — colorspaces.rs:97-101
  float cm[12];    // cmatrix [[f32;4];3] — colorspaces.rs:90-95
  float rgbm[9];   // XYZ_D65_33 = inverse(SRGB_D65_33) in f32 — color_conversions.rs:8
  int use_e;       // 0: the 4th (E) channel is identically 0 and the matrix is finite -> skip its term
  int linear;      // settings.linear: skip gamma (gamma.rs:17-18)
  float one, mone; // 1.0f and -1.0f as run-time values for the packed adds of ipb_fused.cu (see PkAdd there)
  SplineDev sp;
};

// ---------------------------------------------------------------- exact helpers

// x / d for a constant d with rc = RN(1/d): bit-identical to IEEE division (tools/verify_constdiv.c).
__device__ __forceinline__ float div_rc(float x, float d, float rc) {
  float q = x * rc;
  float r = fmaf(-q, d, x);
  return fmaf(r, rc, q);
}
#define IPB_DIVC(x, d) ::ipb::div_rc((x), (d), 1.0f / (d))
// RC = true: the 3-instruction form, exact whenever x and x/d are finite and normal (or zero) — the fused kernels,
// whose launch is gated on parameter bounds that guarantee it (ipb_host.cu fused_params_bounded).  RC = false: IEEE
// division for every input including inf/NaN/denormals — the per-op kernels.
template <bool RC>
__device__ __forceinline__ float divc(float x, float d) {
  return RC ? div_rc(x, d, 1.0f / d) : __fdiv_rn(x, d);
}

// ---------------------------------------------------------------- packed f32x2 arithmetic (sm_100: FMUL2/FFMA2)
// Two pixels per instruction, each half with the exact IEEE round-to-nearest result of the scalar operation.
// ptxas contracts a mul.f32x2 feeding an add.f32x2 into one FFMA2 even under --fmad=false (seen in SASS), which
// would change the rounding; so packed additions are issued as fma(a, one, b) with `one` a kernel parameter the
// assembler cannot see through: a*1 + b rounds exactly like a + b, and two FMAs are never merged.
struct F2 { float x, y; };
__device__ __forceinline__ F2 splat(float c) { return F2{c, c}; }
__device__ __forceinline__ F2 pk_mul(F2 a, F2 b) {
  F2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ F2 pk_fma(F2 a, F2 b, F2 c) {
  F2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};"
      " fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ F2 pk_fma_rm(F2 a, F2 b, F2 c) {  // round towards -inf
  F2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};"
      " fma.rm.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ F2 pk_mul(F2 a, float c) { return pk_mul(a, splat(c)); }
struct PkAdd {
  float one, mone;  // ColorParams::one / mone
  __device__ __forceinline__ F2 add(F2 a, F2 b) const { return pk_fma(a, splat(one), b); }
  __device__ __forceinline__ F2 add(F2 a, float c) const { return pk_fma(a, splat(one), splat(c)); }
  __device__ __forceinline__ F2 sub(F2 a, F2 b) const { return pk_fma(b, splat(mone), a); }  // a - b == (-1*b) + a
  __device__ __forceinline__ F2 add_rm(F2 a, float c) const { return pk_fma_rm(a, splat(one), splat(c)); }
};
// x / d for a constant d, both halves (div_rc above)
__device__ __forceinline__ F2 pk_div_rc(F2 x, float d, float rc) {
  F2 q = pk_mul(x, rc);
  F2 r = pk_fma(q, splat(-d), x);
  return pk_fma(r, splat(rc), q);
}
#define IPB_PK_DIVC(x, d) ::ipb::pk_div_rc((x), (d), 1.0f / (d))

// glibc 2.39 cbrtf (sysdeps/ieee754/flt-32/s_cbrtf.c) restated for x > 1; checked bit-identical to the host libm
// (tests/test_gpu_ops.py::test_lab_transfer_above_one).  Only used by the out-of-table fallback of the Lab transfer
// function (color_conversions.rs:103-104,123), i.e. for XYZ ratios above 1.0.  The double-precision Halley step
// is kept in FP64 (half rate on B200): its rounding to f32 is what makes the result glibc's and not the
// correctly rounded cube root.
__constant__ double kCbrtFac[3] = {1.0, 1.2599210498948731648, 1.5874010519681994748};  // factor[2 + xe % 3], xe >= 0
static __device__ __forceinline__ float cbrt_glibc_gt1(float x) {  // x > 1.0 (so: positive, normal or +inf, never NaN)
  const uint32_t bits = __float_as_uint(x);
  const uint32_t xe = (bits >> 23) - 126u;                            // frexpf exponent, 1..129
  const float xm = __uint_as_float((bits & 0x007fffffu) | 0x3f000000u);  // frexpf mantissa: [0.5, 1)
  const double dxm = (double)xm;
  const float u = (float)(0.492659620528969547 + (0.697570460207922770 - 0.191502161678719066 * dxm) * dxm);
  const float t2 = u * u * u;
  const uint32_t q3 = xe / 3u, m3 = xe - 3u * q3;                     // xe / 3 and xe % 3 (C semantics, xe > 0)
  const double dt2 = (double)t2;
  const float ym = (float)((double)u * (dt2 + 2.0 * dxm) / (2.0 * dt2 + dxm) * kCbrtFac[m3]);
  const float res = ym * __uint_as_float((127u + q3) << 23);          // ldexpf: ym in [0.5, 2), q3 <= 43: exact
  return bits == 0x7f800000u ? x : res;                               // +inf: glibc returns x + x
}

// The analytic branch of XYZ_LAB_TRANSFORM.lookup (color_conversions.rs:102-104,120-124) plus the two table-branch
// inputs the masked fast lerp does not handle (-0.0 and NaN).  Called for values with !in_table().
static __device__ __noinline__ float lab_f_slow(float v) {
  const float e = 216.0f / 24389.0f;
  const float k = 24389.0f / 27.0f;
  if (v > 1.0f) return cbrt_glibc_gt1(v);
  if (v < 0.0f) return __fdiv_rn(k * v + 16.0f, 116.0f);  // v < 0 is never > e
  if (v != v) return v;                                  // NaN takes the table branch: a = NaN
  (void)e;
  return __fdiv_rn(k * 0.0f + 16.0f, 116.0f);            // -0.0: table[0] + 0 * (table[1] - table[0])
}

// ---------------------------------------------------------------- TransformLookup (color_conversions.rs:80-115)

// Table access policies: global memory (unfused per-op kernels) or shared memory (fused kernels).
struct LutGlobal {
  const float2 *t;
  __device__ __forceinline__ float2 at(int key) const { return __ldg(t + key); }
};
struct LutShared {
  uint32_t base;  // shared-window byte address of entry 0
  __device__ __forceinline__ float2 at(int key) const {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + ((uint32_t)key << 3)));
    return v;
  }
};

// lookup() table branch for val in [0,1] (or -0.0): pos = val*max; key = trunc(pos); a = pos - trunc(pos);
// v1 + a*(v2 - v1).  floor == trunc for pos >= 0; FADD.RM with 2^23 yields floor(pos) in the low mantissa bits.
template <class Lut>
__device__ __forceinline__ float lut_lerp(const Lut &lut, float val) {
  float pos = val * kLutMax;
  float tf = __fadd_rd(pos, 8388608.0f);
  float base = tf - 8388608.0f;
  float a = pos - base;
  int key = __float_as_int(tf) & 0x3fff;
  float2 e = lut.at(key);
  return e.x + a * e.y;
}

// XYZ_LAB_TRANSFORM.lookup — color_conversions.rs:120-124 with the analytic fallback outside [0,1]
template <class Lut>
__device__ __forceinline__ float lab_f(const Lut &lut, float v) {
  if (v < 0.0f || v > 1.0f) return lab_f_slow(v);
  return lut_lerp(lut, v);
}

// ---------------------------------------------------------------- colour chain

// camera_to_lab + xyz_to_lab — color_conversions.rs:42-55,156-169
template <bool RC, class Lut>
__device__ __forceinline__ void camera_to_lab(const ColorParams &P, const Lut &lab, float r, float g, float b,
                                              float e, float &ol, float &oa, float &ob) {
  r = fminf(r * P.mul[0], 1.0f);
  g = fminf(g * P.mul[1], 1.0f);
  b = fminf(b * P.mul[2], 1.0f);
  float x = r * P.cm[0] + g * P.cm[1] + b * P.cm[2];
  float y = r * P.cm[4] + g * P.cm[5] + b * P.cm[6];
  float z = r * P.cm[8] + g * P.cm[9] + b * P.cm[10];
  if (P.use_e) {
    e = fminf(e * P.mul[3], 1.0f);
    x = x + e * P.cm[3];
    y = y + e * P.cm[7];
    z = z + e * P.cm[11];
  }
  float xr = divc<RC>(x, 0.95047f);
  float yr = y;  // y / 1.0
  float zr = divc<RC>(z, 1.08883f);
  float fx = lab_f(lab, xr);
  float fy = lab_f(lab, yr);
  float fz = lab_f(lab, zr);
  float l = 116.0f * fy - 16.0f;
  float a = 500.0f * (fx - fy);
  float bb = 200.0f * (fy - fz);
  ol = divc<RC>(l, 100.0f);
  oa = divc<RC>(a + 127.0f, 255.0f);
  ob = divc<RC>(bb + 127.0f, 255.0f);
}

// SplineFunc::interpolate — curves.rs:126-157, statement by statement, including the binary search over
// points[0..nseg): for sorted knots it ends at the last knot below val (or returns y exactly on a knot); for unsorted
// knots and NaN it does whatever the reference's search does (NaN fails every comparison and returns y[(nseg-1)/2]).
// Used by the per-op kernels; the fused kernels take sorted knots only (ipb_fused.cu spline_eval_smem).
__device__ __forceinline__ float spline_eval(const SplineDev &s, float val) {
  const int last = s.n - 1;
  if (val >= s.x[last]) return s.y[last];
  if (val <= s.x[0]) return s.y[0];
  int low = 0, high = s.nseg - 1;
  while (low <= high) {
    const int mid = (low + high) / 2;
    const float xhere = s.x[mid];
    if (xhere < val) low = mid + 1;
    else if (xhere > val) high = mid - 1;
    else return s.y[mid];
  }
  const int i = high > 0 ? high : 0;
  const float diff = val - s.x[i];
  return s.y[i] + s.c1[i] * diff + s.c2[i] * diff * diff + s.c3[i] * diff * diff * diff;
}

// lab_to_xyz + lab_to_rgb — color_conversions.rs:58-65,172-191
template <bool RC>
__device__ __forceinline__ void lab_to_rgb(const ColorParams &P, float l, float a, float b, float &r, float &g,
                                           float &bl) {
  const float e = 216.0f / 24389.0f;
  const float k = 24389.0f / 27.0f;
  float cl = l * 100.0f;
  float ca = (a * 255.0f) - 127.0f;
  float cb = (b * 255.0f) - 127.0f;
  float fy = divc<RC>(cl + 16.0f, 116.0f);
  float fx = divc<RC>(ca, 500.0f) + fy;
  float fz = fy - divc<RC>(cb, 200.0f);
  float fx3 = fx * fx * fx;
  float xr = fx3 > e ? fx3 : divc<RC>(116.0f * fx - 16.0f, k);
  float yr = cl > k * e ? fy * fy * fy : divc<RC>(cl, k);
  float fz3 = fz * fz * fz;
  float zr = fz3 > e ? fz3 : divc<RC>(116.0f * fz - 16.0f, k);
  float x = xr * 0.95047f;
  float y = yr;  // * 1.0
  float z = zr * 1.08883f;
  r = x * P.rgbm[0] + y * P.rgbm[1] + z * P.rgbm[2];
  g = x * P.rgbm[3] + y * P.rgbm[4] + z * P.rgbm[5];
  bl = x * P.rgbm[6] + y * P.rgbm[7] + z * P.rgbm[8];
}

// OpGamma per element — gamma.rs:21: apply_srgb_gamma(x.max(0.0).min(1.0)); the clamp keeps it on the table
template <class Lut>
__device__ __forceinline__ float gamma_elem(const Lut &gam, float v) {
  return lut_lerp(gam, fminf(fmaxf(v, 0.0f), 1.0f));
}

// demosaiced RGBE -> final RGB (to_lab, basecurve, from_lab, gamma): the whole chain for one pixel
template <bool RC, class Lut>
__device__ __forceinline__ void color_chain(const ColorParams &P, const Lut &lab, const Lut &gam, float r, float g,
                                            float b, float e, float &or_, float &og, float &ob) {
  float l, a, bb;
  camera_to_lab<RC>(P, lab, r, g, b, e, l, a, bb);
  if (P.sp.n > 0) l = spline_eval(P.sp, l);
  lab_to_rgb<RC>(P, l, a, bb, or_, og, ob);
  if (!P.linear) {
    or_ = gamma_elem(gam, or_);
    og = gamma_elem(gam, og);
    ob = gamma_elem(gam, ob);
  }
}

// output8bit — color_conversions.rs:323-325: (v*256).max(0).min(255) as u8.  FADD.RZ with 2^23 leaves
// trunc() in the low mantissa byte (the clamp keeps the value in [0,255]).
__device__ __forceinline__ uint32_t output8bit(float v) {
  float t = fminf(fmaxf(v * 256.0f, 0.0f), 255.0f);
  return __float_as_uint(__fadd_rz(t, 8388608.0f)) & 0xffu;
}
// output16bit — color_conversions.rs:328-330: (v*65535).round().max(0).min(65535) as u16
__device__ __forceinline__ uint32_t output16bit(float v) {
  float t = fminf(fmaxf(roundf(v * 65535.0f), 0.0f), 65535.0f);
  return (uint32_t)t;
}

// gofloat level mapping — gofloat.rs:127: ((v as f32 - black) / range).min(1.0).  With exact != 0 the host
// has verified (all 65536 inputs) that the 3-instruction division is bit-identical for this black/range.
__device__ __forceinline__ float golevel(float v, float black, float range, float rc, int exact_rc) {
  float num = v - black;
  float q = exact_rc ? div_rc(num, range, rc) : __fdiv_rn(num, range);
  return fminf(q, 1.0f);
}

}  // namespace ipb
