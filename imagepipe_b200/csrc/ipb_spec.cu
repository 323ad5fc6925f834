// ipb_spec.cu — the speculative 8-bit raw -> sRGB kernels: k_spec8 (full resolution: the roofline kernel of BASELINE
// configs 2, 3 and 5) and k_spec8_scaled (behind scaled_demosaic: config 4, further down).
//
// output_8bit of a full-resolution RGB Bayer frame is a function  u16 CFA samples -> 3 bytes per pixel.  The bytes
// are decided by which of the 255 thresholds of  v -> output8bit(apply_srgb_gamma(clamp(v)))  (gamma.rs:21,
// color_conversions.rs:323-325; monotone in v, ipb_host.cu build_gamma8) the linear value v of a channel lies above.
// So the reference's f32 arithmetic has to be reproduced bit for bit only where v is close to a threshold:
//
//   cheap pass   every interior pixel goes through a restatement of the chain in contracted arithmetic: level
//                mapping by a reciprocal multiply, Bayer demosaic on pixel pairs (packed f32x2), white balance and both
//                matrices folded into FFMA2 chains, the Lab transfer function on the XU pipe (ex2(lg2(v)/3), no table),
//                to_lab -> basecurve -> from_lab collapsed to   t = f(v) + (S(f(Y)) - f(Y)),  v' = t^3   with S the
//                basecurve in f-space read from a slope/intercept table, and the 8-bit gamma through a 32-bit
//                {byte, threshold} table in fixed point.  Its linear value differs from the reference's by at most
//                delta (derivation: DESIGN.md 3.1 "The bound"; computed per parameter set on the host from the actual
//                matrices and curve, ipb_spec_host.cu spec_build; measured by ipb_pipeline_spec_probe).
//   certificate  a channel whose cheap value is farther than delta from every threshold has, by monotonicity, exactly
//                the reference's byte.  One add and one mask per channel produce byte and distance together.
//   fix-up       pixels with a channel inside +-delta of a threshold (about 2 %), and the frame's border pixels, are
//                queued in shared memory per tile and recomputed by the bit-exact code of ipb_device.cuh (same
//                arithmetic as the per-op kernels and k_fused_full) from the tile's level-mapped samples, which are
//                still in shared memory: after the tile's barrier the first ceil(n / 32) warps take the n entries,
//                the other warps go on with the next tile; the exact bytes are stored over the cheap ones.
//
// Result: byte-identical to k_fused_full / the oracle by construction, at about a third of the instructions.
// Tile pipeline as in k_fused_full: persistent CTAs, raw u16 boxes by TMA (cp.async.bulk.tensor.2d + mbarrier) one
// tile ahead, conversion into a double-buffered f32 tile whose even and odd columns live in separate planes (so that
// the two same-kind pixels of a four-pixel task sit in one aligned register pair and no window load has a bank
// conflict), one __syncthreads per tile plus an immediate second one that frees the queue.  Three-colour patterns other
// than RGB Bayer (X-Trans ...) take MODE 4: row-major tile, per-position tap masks, exact means, the same chain.  A
// batch of frames of one geometry (BATCH) is one launch whose tile sequence runs through the frames.
#include <cuda.h>

#include "ipb_internal.h"
#include "ipb_scaled.cuh"
#include "ipb_spec.h"

namespace ipb {

namespace {

constexpr int kTW = 128, kTH = 32;
constexpr int kTileStride = kTW + 16;  // TMA box columns: frame col tx0-8 .. tx0+kTW+7
constexpr int kTileRows = kTH + 2;
constexpr int kTileElems = kTileRows * kTileStride;
constexpr int kStageElems = (kTileElems * 2 + 127) / 128 * 64;
constexpr int kPS = kTileStride / 2;   // plane stride (floats): 72
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kG8Offset = 32768 - (int)kSpecSmemBase;  // the gamma table starts on a 32 KB boundary of the shared window
constexpr int kSplRows = kMaxSplinePts + 2;
constexpr int kChunks = (kTileElems / 8 + 31) / 32;  // conversion chunks per tile: 32 lanes x 8 samples

struct SmemFront {
  float2 stab[kSpecSTabEntries];                // basecurve in f-space: {intercept, slope} per segment
  alignas(16) float spl[kSplRows][8];           // exact basecurve: [0] below the first knot, [1 + i] segment i, [n] at / above the last
  float thr[260];                               // the 255 thresholds of output8bit(apply_srgb_gamma(v)), then +inf
  uint2 taps[144];                              // generic patterns: per position the 9-bit tap masks of colours 0, 1 (.x) and 2 (.y)
  uint8_t pat[144];                             // ... and one period of the colour table (ph rows of pw)
  alignas(16) SpecParams sp;                    // copies for the out-of-line exact path (a generic pointer into the
  alignas(16) ColorParams cp;                   // constant bank would turn every parameter access into a global load)
  alignas(8) unsigned long long mbar;
  alignas(8) unsigned long long mbar_tab;
  int conv_ctr[2];
  int qn[2];                                    // queue length of the tile being computed / the tile being recomputed
};
constexpr int kQueueCap = kTW * kTH;            // every pixel of a tile: the queue cannot overflow
static_assert(kG8Offset - (int)sizeof(SmemFront) - kQueueCap * 2 >= 0, "front part of the shared window is full");

struct SmemSpec : SmemFront {
  uint16_t queue[kQueueCap];                    // uncertified pixels of the current tile: row * kTW + column inside the tile
  unsigned char fill[kG8Offset - (int)sizeof(SmemFront) - kQueueCap * 2];
  uint32_t g8a[kSpecG8Entries];                 // {byte, threshold} fixed-point gamma table; 32 KB aligned: its
                                                // entries are addressed by (bits & 0x7ffc) | base, one LOP3
  float plane[2][kTileRows][2][kPS];            // [buffer][tile row][even / odd columns][column / 2]
  alignas(128) uint16_t raw[kStageElems];       // TMA destination
};
static_assert(offsetof(SmemSpec, g8a) == kG8Offset, "gamma table must sit on a 32 KB boundary of the shared window");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
      "SPEC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra SPEC_DONE;\n"
      "bra SPEC_WAIT;\n"
      "SPEC_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---------------------------------------------------------------- contracted packed arithmetic (cheap pass only)
__device__ __forceinline__ F2 a2(F2 a, F2 b) {
  F2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ F2 m2(F2 a, F2 b) { return pk_mul(a, b); }
__device__ __forceinline__ F2 f2(F2 a, F2 b, F2 c) { return pk_fma(a, b, c); }
__device__ __forceinline__ float lg2a(float v) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float ex2a(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ F2 lds64f(uint32_t addr) {
  F2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

constexpr float kLabE = 216.0f / 24389.0f;   // color_conversions.rs:121
constexpr float kLabK = 24389.0f / 27.0f;    // :122

// Lab transfer function (color_conversions.rs:120-124) of both halves: cube root on the XU pipe above e, the line below
__device__ __forceinline__ F2 lab_f_cheap(F2 v) {
  const F2 lg = m2(F2{lg2a(v.x), lg2a(v.y)}, splat(1.0f / 3.0f));
  const F2 lin = f2(v, splat(kLabK / 116.0f), splat(16.0f / 116.0f));
  return F2{v.x > kLabE ? ex2a(lg.x) : lin.x, v.y > kLabE ? ex2a(lg.y) : lin.y};
}
// its inverse (color_conversions.rs:172-191): t^3 above cbrt(e) = 6/29, the line below
__device__ __forceinline__ F2 lab_g_cheap(F2 t) {
  const F2 t2 = m2(t, t);
  const F2 lin = f2(t, splat(116.0f / kLabK), splat(-16.0f / kLabK));
  const float t0 = 6.0f / 29.0f;
  return F2{t.x > t0 ? t2.x * t.x : lin.x, t.y > t0 ? t2.y * t.y : lin.y};
}
// basecurve in f-space, S(fy) = (100*spline((116*fy - 16)/100) + 16)/116, from the slope/intercept table
__device__ __forceinline__ float s_cheap(uint32_t stab_bias, float fy, float tf) {
  const F2 e = lds64f((__float_as_uint(tf) << 3) + stab_bias);
  return fmaf(fy, e.y, e.x);
}

// The colour chain of two pixels in the cheap arithmetic.  In: demosaiced camera RGB (white balance not applied).
// Out: per channel the sum  table entry + fixed-point value  whose top byte is the 8-bit result and whose low 24
// bits are the distance to the threshold (+ delta), and with PROBE the clamped linear values themselves.
template <bool PROBE>
__device__ __forceinline__ float chain_pair(const SpecParams &p, uint32_t g8_base, uint32_t stab_bias, F2 cr, F2 cg, F2 cb,
                                            uint32_t s[6], float lin[6]) {
  // white balance + clip (colorspaces.rs:97-101, color_conversions.rs:43-46) as a clip against 1/mul; the gains are
  // folded into the matrix.  mul[1] is 1 by construction and demosaiced values never exceed 1: green needs no clip.
  cr = F2{fminf(cr.x, p.lim_r), fminf(cr.y, p.lim_r)};
  cb = F2{fminf(cb.x, p.lim_b), fminf(cb.y, p.lim_b)};
  // camera -> XYZ ratios (matrix rows divided by the white point)
  const F2 xr = f2(cb, splat(p.m[0][2]), f2(cg, splat(p.m[0][1]), m2(cr, splat(p.m[0][0]))));
  const F2 yr = f2(cb, splat(p.m[1][2]), f2(cg, splat(p.m[1][1]), m2(cr, splat(p.m[1][0]))));
  const F2 zr = f2(cb, splat(p.m[2][2]), f2(cg, splat(p.m[2][1]), m2(cr, splat(p.m[2][0]))));
  const F2 fx = lab_f_cheap(xr), fy = lab_f_cheap(yr), fz = lab_f_cheap(zr);
  // basecurve: fy' = S(fy); a and b are untouched, so fx' = fx + (fy' - fy), fz' = fz + (fy' - fy)
  const float ux = __saturatef(fmaf(fy.x, p.s_scale, p.s_off)), uy = __saturatef(fmaf(fy.y, p.s_scale, p.s_off));
  const F2 tf = pk_fma_rm(F2{ux, uy}, splat((float)(kSpecSTabN + 2)), splat(8388608.0f));
  const F2 fyp{s_cheap(stab_bias, fy.x, tf.x), s_cheap(stab_bias, fy.y, tf.y)};
  const F2 d = f2(fy, splat(-1.0f), fyp);
  const F2 X = lab_g_cheap(a2(fx, d)), Y = lab_g_cheap(fyp), Z = lab_g_cheap(a2(fz, d));
  // XYZ ratios -> linear sRGB (matrix columns multiplied by the white point), clamp to [0, 1] in the last FMA
  const F2 r0 = f2(Y, splat(p.ro[0][1]), m2(X, splat(p.ro[0][0])));
  const F2 g0 = f2(Y, splat(p.ro[1][1]), m2(X, splat(p.ro[1][0])));
  const F2 b0 = f2(Y, splat(p.ro[2][1]), m2(X, splat(p.ro[2][0])));
  const F2 r{__saturatef(fmaf(Z.x, p.ro[0][2], r0.x)), __saturatef(fmaf(Z.y, p.ro[0][2], r0.y))};
  const F2 g{__saturatef(fmaf(Z.x, p.ro[1][2], g0.x)), __saturatef(fmaf(Z.y, p.ro[1][2], g0.y))};
  const F2 b{__saturatef(fmaf(Z.x, p.ro[2][2], b0.x)), __saturatef(fmaf(Z.y, p.ro[2][2], b0.y))};
  if (PROBE) { lin[0] = r.x; lin[1] = r.y; lin[2] = g.x; lin[3] = g.y; lin[4] = b.x; lin[5] = b.y; }
  // fixed point: u = one[c] + v*(1 - 2^-13) has the bit pattern 0x3F800000 + F + deltaF_c, F = round(v * (2^23 - 2^10));
  // table segment = floor(v * 8191), taken from a second float whose mantissa is floor(v * 32764): bits 2..14
  const float c = 1.0f - 1.0f / 8192.0f, c1 = 32764.0f / 8388608.0f;
  const F2 ur = f2(r, splat(c), splat(p.one[0])), ug = f2(g, splat(c), splat(p.one[1])), ub = f2(b, splat(c), splat(p.one[2]));
  const F2 kr = pk_fma_rm(r, splat(c1), splat(1.0f)), kg = pk_fma_rm(g, splat(c1), splat(1.0f)), kb = pk_fma_rm(b, splat(c1), splat(1.0f));
  const float uu[6] = {ur.x, ur.y, ug.x, ug.y, ub.x, ub.y}, kk[6] = {kr.x, kr.y, kg.x, kg.y, kb.x, kb.y};
#pragma unroll
  for (int i = 0; i < 6; i++) s[i] = lds32u((__float_as_uint(kk[i]) & 0x7ffcu) | g8_base) + __float_as_uint(uu[i]);
  return fminf(yr.x, yr.y);  // the caller checks the certified domain (Y ratio >= kSpecYMin)
}

struct Window {
  F2 En, Ec, Es, On, Oc, Os;  // columns (x0, x0+2) and (x0+1, x0+3) of rows y-1, y, y+1
  float e2n, e2c, e2s;        // column x0+4
  float omn, omc, oms;        // column x0-1
};

// demosaic::full for an interior four-pixel task of an RGB Bayer frame (demosaic.rs:67-119: bilinear means, the
// centre's own colour passed through).  GF: the row starts with green (x0 is even).  Pixels (0,2) and (1,3) are of the
// same kind and travel as packed pairs.  Out: (row colour a, green, other colour o) of both pairs.
template <bool GF>
__device__ __forceinline__ void demosaic_pairs(const Window &w, F2 &a02, F2 &g02, F2 &o02, F2 &a13, F2 &g13, F2 &o13) {
  const F2 q{0.25f, 0.25f}, h{0.5f, 0.5f};
  if (!GF) {
    // pixels 0, 2 on the row's colour
    const F2 W{w.omc, w.Oc.x}, NW{w.omn, w.On.x}, SW{w.oms, w.Os.x};
    a02 = w.Ec;
    g02 = m2(a2(a2(a2(w.En, W), w.Oc), w.Es), q);
    o02 = m2(a2(a2(a2(NW, w.On), SW), w.Os), q);
    // pixels 1, 3 on green
    const F2 E{w.Ec.y, w.e2c};
    g13 = w.Oc;
    a13 = m2(a2(w.Ec, E), h);
    o13 = m2(a2(w.On, w.Os), h);
  } else {
    const F2 W{w.omc, w.Oc.x};
    g02 = w.Ec;
    a02 = m2(a2(W, w.Oc), h);
    o02 = m2(a2(w.En, w.Es), h);
    const F2 E{w.Ec.y, w.e2c}, NE{w.En.y, w.e2n}, SE{w.Es.y, w.e2s};
    a13 = w.Oc;
    g13 = m2(a2(a2(a2(w.On, w.Ec), E), w.Os), q);
    o13 = m2(a2(a2(a2(w.En, NE), w.Es), SE), q);
  }
}

// ---------------------------------------------------------------- exact pixel (fix-up)
// SplineFunc::interpolate (curves.rs:126-157) from the table in shared memory — the evaluation k_fused_full uses
// (ipb_fused.cu spline_eval_smem): entry = number of knots <= val; entries 0 and n are the constant end pieces
// {y, 0, 0, 0}.  The launch is gated on finite, strictly increasing knots (ipb_host.cu fused_params_bounded).
__device__ __forceinline__ float spline_exact(const float (*spl)[8], const SplineDev &s, float val) {
  int idx = (val >= s.x[0] ? 1 : 0) + (val >= s.x[1] ? 1 : 0);
  for (int j = 2; j < s.n; j++) idx += val >= s.x[j] ? 1 : 0;
  const float4 c = *reinterpret_cast<const float4 *>(spl[idx]);
  const float c3 = spl[idx][4];
  const float diff = val - c.x;
  return c.y + c.z * diff + c.w * diff * diff + c3 * diff * diff * diff;
}

// XYZ_LAB_TRANSFORM.lookup (color_conversions.rs:102-114,120-124) without divergent calls, the scheme of k_fused_full
// (ipb_fused.cu lab_outside_table): the table lerp for +0 <= v <= 1, the host libm's cbrtf from the context's table for
// 1 < v <= 1.5, the line for v < 0 (its division in the verified reciprocal form), lab_f_slow for what is left
// (v > 1.5, -0.0, NaN).
__device__ __forceinline__ float lab_f_exact(const float2 *__restrict__ lut, const float *__restrict__ cbrt_tab, float v) {
  const float pos = v * kLutMax;
  const float tf = __fadd_rd(pos, 8388608.0f);
  const float a = pos - (tf - 8388608.0f);
  const float2 e = __ldg(lut + (__float_as_uint(tf) & 0x1fffu));
  float r = e.x + a * e.y;
  const uint32_t u = __float_as_uint(v), first = 0x3f800001u, size = 1u << 22;
  if (u - first < size) r = __ldg(cbrt_tab + (u - first));
  if (v < 0.0f) r = div_rc(kLabK * v + 16.0f, 116.0f, 1.0f / 116.0f);
  if (u - (first + size) <= 0x80000000u - (first + size) || u > 0xff800000u) r = lab_f_slow(v);
  return r;
}

// demosaic::full (demosaic.rs:67-119) for ONE pixel of an RGB Bayer frame from its nine level-mapped taps t[0..8] (raster
// order; whatever sits in a tap outside the frame is ignored), in the reference's arithmetic.  Any position, frame
// borders included (a tap outside the frame is dropped from sum and count, :103-107).  Taps reach their colour's sum
// in the reference's raster order.  Both kinds of site run the same instructions (four means, a select), so a warp of
// queue entries does not diverge.  `phase` holds the colour of position (row & 1, col & 1) in bits
// 2*(2*(row&1)+(col&1)).
__device__ __forceinline__ void exact_rgb_bayer(const SpecParams &p, uint32_t phase, int x, int y, const float t[9], float &r,
                                                float &g, float &b) {
  const bool hn = y > 0, hs = y < p.height - 1, hw = x > 0, he = x < p.width - 1;
  const float v = t[4];
  const int c = (phase >> (2 * (2 * (y & 1) + (x & 1)))) & 3;            // this site's colour
  const int ch = (phase >> (2 * (2 * (y & 1) + ((x & 1) ^ 1)))) & 3;     // the colour of its left / right neighbours
  const bool gsite = c == 1;
  const int first = gsite ? ch : c;  // colour (0 or 2) that receives `a0`
  if (hn && hs && hw && he) {
    // all nine taps exist (nearly every recomputed pixel): the same sums in the same order, divisions by 4 and 2 as
    // exact scalings
    const float mg = ((((0.0f + t[1]) + t[3]) + t[5]) + t[7]) * 0.25f, md = ((((0.0f + t[0]) + t[2]) + t[6]) + t[8]) * 0.25f;
    const float mh = ((0.0f + t[3]) + t[5]) * 0.5f, mv = ((0.0f + t[1]) + t[7]) * 0.5f;
    const float a0 = gsite ? mh : v, a1 = gsite ? mv : md;
    r = first == 0 ? a0 : a1;
    g = gsite ? v : mg;
    b = first == 0 ? a1 : a0;
    return;
  }
  auto mean = [](float s, int n) { return n == 4 ? s * 0.25f : n == 2 ? s * 0.5f : n ? __fdiv_rn(s, (float)n) : 0.0f; };
  // sums start at +0.0 and skip missing taps (x + 0.0 == x for these sums, which are never -0.0)
  float sg = 0.0f, sd = 0.0f, sh = 0.0f, sv = 0.0f;
  if (hn && hw) sd = sd + t[0];
  if (hn) { sg = sg + t[1]; sv = sv + t[1]; }
  if (hn && he) sd = sd + t[2];
  if (hw) { sg = sg + t[3]; sh = sh + t[3]; }
  if (he) { sg = sg + t[5]; sh = sh + t[5]; }
  if (hs && hw) sd = sd + t[6];
  if (hs) { sg = sg + t[7]; sv = sv + t[7]; }
  if (hs && he) sd = sd + t[8];
  const int nv = (int)hn + (int)hs, nh = (int)hw + (int)he;
  const float mg = mean(sg, nv + nh), md = mean(sd, nv * nh), mh = mean(sh, nh), mv = mean(sv, nv);
  // green site: own sample, left/right mean for colour ch, up/down mean for the third; red / blue site: own sample,
  // edge mean for green, corner mean for the third
  const float a0 = gsite ? mh : v, a1 = gsite ? mv : md;
  r = first == 0 ? a0 : a1;
  g = gsite ? v : mg;
  b = first == 0 ? a1 : a0;
}

// The same for any three-colour pattern (X-Trans ...): `pat` is one period of the colour table (ph rows of pw bytes).
// Statement for statement demosaic.rs:77-116: a tap of the centre's own colour other than the centre itself is discarded
// (:87), a tap outside the frame is skipped (:103-104), a colour's value is sum / count, or 0.0 without taps (:110-114).
__device__ __forceinline__ void exact_rgb_generic(const SpecParams &p, const uint8_t *pat, int x, int y, const float t[9], float &r,
                                                  float &g, float &b) {
  const int pr = y % p.ph, pc = x % p.pw;
  const int pix = pat[pr * p.pw + pc];
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
  int n0 = 0, n1 = 0, n2 = 0;
#pragma unroll
  for (int dy = -1; dy <= 1; dy++)
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
      const int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= p.height || xx < 0 || xx >= p.width) continue;
      const int rr = pr + dy < 0 ? p.ph - 1 : (pr + dy == p.ph ? 0 : pr + dy), cc = pc + dx < 0 ? p.pw - 1 : (pc + dx == p.pw ? 0 : pc + dx);
      const int oc = pat[rr * p.pw + cc];
      if (oc == pix && (dx != 0 || dy != 0)) continue;
      const float v = t[(dy + 1) * 3 + dx + 1];
      if (oc == 0) { s0 = s0 + v; n0++; }
      else if (oc == 1) { s1 = s1 + v; n1++; }
      else if (oc == 2) { s2 = s2 + v; n2++; }
    }
  r = n0 ? __fdiv_rn(s0, (float)n0) : 0.0f;
  g = n1 ? __fdiv_rn(s1, (float)n1) : 0.0f;
  b = n2 ? __fdiv_rn(s2, (float)n2) : 0.0f;
}

// to_lab + basecurve + from_lab of one demosaiced pixel in the reference's arithmetic: the per-pixel code of
// ipb_device.cuh with the verified reciprocal divisions and k_fused_full's table for cube roots above one.
// Out: linear RGB before OpGamma.
__device__ __forceinline__ void exact_chain(const SpecParams &p, const ColorParams &P, const float (*spl)[8], float r, float g,
                                            float b, float out[3]) {
  // camera_to_lab (color_conversions.rs:42-55,156-169), statement for statement as ipb_device.cuh camera_to_lab<true>
  const float cr = fminf(r * P.mul[0], 1.0f), cg = fminf(g * P.mul[1], 1.0f), cb = fminf(b * P.mul[2], 1.0f);
  const float X = cr * P.cm[0] + cg * P.cm[1] + cb * P.cm[2];
  const float Y = cr * P.cm[4] + cg * P.cm[5] + cb * P.cm[6];
  const float Z = cr * P.cm[8] + cg * P.cm[9] + cb * P.cm[10];
  const float fx = lab_f_exact(p.lut_lab, p.cbrt_tab, divc<true>(X, 0.95047f));
  const float fy = lab_f_exact(p.lut_lab, p.cbrt_tab, Y);
  const float fz = lab_f_exact(p.lut_lab, p.cbrt_tab, divc<true>(Z, 1.08883f));
  float l = divc<true>(116.0f * fy - 16.0f, 100.0f);
  const float a = divc<true>(500.0f * (fx - fy) + 127.0f, 255.0f);
  const float bb = divc<true>(200.0f * (fy - fz) + 127.0f, 255.0f);
  if (P.sp.n > 0) l = spline_exact(spl, P.sp, l);
  lab_to_rgb<true>(P, l, a, bb, out[0], out[1], out[2]);
}

// gofloat (gofloat.rs:127) of the nine taps of pixel (x, y) straight from the raw frame, one 2-byte load per tap (probe)
__device__ __forceinline__ void taps_from_frame(const SpecParams &p, int x, int y, float t[9]) {
  const uint16_t *ctr = p.raw + (long long)(y + p.crop_y - p.src_row0) * p.raw_pitch + p.crop_x + x;
  const bool hn = y > 0, hs = y < p.height - 1, hw = x > 0, he = x < p.width - 1;
#pragma unroll
  for (int dy = -1; dy <= 1; dy++)
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
      const bool have = (dy < 0 ? hn : dy > 0 ? hs : true) && (dx < 0 ? hw : dx > 0 ? he : true);
      const float raw = have ? (float)__ldg(ctr + dy * p.raw_pitch + dx) : 0.0f;
      t[(dy + 1) * 3 + dx + 1] = fminf(div_rc(raw - p.black, p.range, p.range_rc), 1.0f);
    }
}

// The nine level-mapped taps of tile pixel (row r, column c) from the tile's f32 planes in shared memory — the very values
// the cheap pass read (gofloat.rs:127 applied once per sample by convert_tile).  Tile row r + 1 is frame row ty0 + r, tile
// column c + 8 is frame column tx0 + c.  ROWMAJOR: the generic-pattern layout; otherwise even / odd columns in two planes.
template <bool ROWMAJOR>
__device__ __forceinline__ void taps_from_tile(uint32_t tile_base, int r, int c, float t[9]) {
#pragma unroll
  for (int dy = 0; dy < 3; dy++)
#pragma unroll
    for (int dx = 0; dx < 3; dx++) {
      const int tc = c + 7 + dx;
      const int idx = ROWMAJOR ? (r + dy) * kTileStride + tc : (r + dy) * (2 * kPS) + (tc & 1) * kPS + (tc >> 1);
      t[dy * 3 + dx] = lds32f(tile_base + (uint32_t)idx * 4u);
    }
}

// output8bit(apply_srgb_gamma(clamp(v))) (gamma.rs:21, color_conversions.rs:323-325) exactly, from shared memory only:
// the byte is the number of thresholds T[j] <= v (ipb_host.cu build_gamma8: the function is a verified step function
// of v).  The fixed-point table of the cheap pass names it to within one — its segment holds at most one threshold and
// its integer comparison can only be off where v is within a unit of that threshold — and two exact comparisons settle it.
__device__ __forceinline__ uint32_t gamma8_exact(uint32_t g8_base, uint32_t thr_base, float v) {
  const float vc = fminf(fmaxf(v, 0.0f), 1.0f);
  const float u = fmaf(vc, 1.0f - 1.0f / 8192.0f, 1.0f), k = __fmaf_rd(vc, 32764.0f / 8388608.0f, 1.0f);
  const uint32_t guess = (lds32u((__float_as_uint(k) & 0x7ffcu) | g8_base) + __float_as_uint(u)) >> 24;
  const uint32_t lo = guess > 0u ? guess - 1u : 0u;
  const float t0 = lds32f(thr_base + lo * 4u), t1 = lds32f(thr_base + lo * 4u + 4u);
  return lo + (vc >= t0 ? 1u : 0u) + (vc >= t1 ? 1u : 0u);
}

// {1/n rounded to nearest, n} for tap counts n = 0..9 (entry 0 divides the empty sum by 1: +0.0)
__constant__ float2 kTapRcpS[10] = {{0.0f, 1.0f}, {1.0f, 1.0f}, {0.5f, 2.0f}, {1.0f / 3.0f, 3.0f}, {0.25f, 4.0f},
                                    {1.0f / 5.0f, 5.0f}, {1.0f / 6.0f, 6.0f}, {1.0f / 7.0f, 7.0f}, {0.125f, 8.0f},
                                    {1.0f / 9.0f, 9.0f}};
// One colour of demosaic::full for an interior pixel of any pattern, exactly as the reference rounds it: the selected taps
// summed in raster order (predicated adds), s / n through the three-instruction reciprocal form, which equals IEEE
// division for every divisor 1..9 (tools/verify_constdiv.c) — the scheme of k_fused_full (ipb_fused.cu bin_mean_rc).
__device__ __forceinline__ float bin_mean_exact(uint32_t m, const float v[9]) {
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; i++)
    if ((m >> i) & 1u) s = s + v[i];
  const float2 e = kTapRcpS[__popc(m)];
  return div_rc(s, e.y, e.x);
}

// queue entry: row * kTW + column inside the tile whose planes start at tile_base and whose first pixel is (tx0, ty0).
// Out of line: the exact path keeps its registers (and the instruction cache footprint of its ~400 instructions) to
// itself; p and P are the copies in shared memory.
__device__ __noinline__ void fixup_entry(const SpecParams &p, const ColorParams &P, const float (*spl)[8], uint32_t phase,
                                         const uint8_t *pat, const uint2 *taps, uint32_t g8_base, uint32_t thr_base, uint32_t tile_base, int tx0,
                                         int ty0, uint8_t *out_f, uint32_t entry) {
  const int r = (int)(entry / kTW), c = (int)(entry % kTW), x = tx0 + c, y = ty0 + r;
  float t[9], v[3], cr, cg, cb;
  if (pat) {   // uniform: one pattern kind per launch
    taps_from_tile<true>(tile_base, r, c, t);
    if (x >= 1 && x <= p.width - 2 && y >= 1 && y <= p.height - 2) {
      // all nine taps exist: the tap masks of the pixel's pattern position and the exact means of the cheap pass
      const int pr = y - (int)__umulhi((uint32_t)y, p.rcp_ph) * p.ph, pc = x - (int)__umulhi((uint32_t)x, p.rcp_pw) * p.pw;
      const uint2 m = taps[pr * p.pw + pc];
      cr = bin_mean_exact(m.x & 0xffffu, t);
      cg = bin_mean_exact(m.x >> 16, t);
      cb = bin_mean_exact(m.y & 0xffffu, t);
    } else {
      exact_rgb_generic(p, pat, x, y, t, cr, cg, cb);
    }
  } else {
    taps_from_tile<false>(tile_base, r, c, t);
    exact_rgb_bayer(p, phase, x, y, t, cr, cg, cb);
  }
  exact_chain(p, P, spl, cr, cg, cb, v);
  uint8_t *o = out_f + ((size_t)(y - p.out_row0) * (size_t)p.width + (size_t)x) * 3;
  o[0] = (uint8_t)gamma8_exact(g8_base, thr_base, v[0]);
  o[1] = (uint8_t)gamma8_exact(g8_base, thr_base, v[1]);
  o[2] = (uint8_t)gamma8_exact(g8_base, thr_base, v[2]);
}

// position of a tile: frame of the batch, tile row and column inside the frame
struct TilePos { int f, tyi, txi; };
__device__ __forceinline__ void issue_tile(const SpecParams &p, const CUtensorMap *tmap, uint32_t raw_stage, uint32_t bar,
                                           const TilePos &q) {
  const int x = q.txi * kTW - 8 + p.crop_x;
  const int y = p.out_row0 + q.tyi * kTH - 1 + p.crop_y - p.src_row0 + q.f * p.frame_src_rows;   // q.f == 0 without a batch
  mbar_expect_tx(bar, kTileElems * (uint32_t)sizeof(uint16_t));
  tma_load_2d(raw_stage, tmap, x, y, bar);
}

// the twelve channel sums of a task, in pixel order (r, g, b of pixel 0 .. 3) -> the three output words and the mask of
// pixels whose certificate failed (a channel within deltaF_c of a threshold) or that lie outside the certified domain
__device__ __forceinline__ uint32_t pack_and_certify(const SpecParams &p, const uint32_t c[12], float ya, float yb, uint32_t ya_mask,
                                                     uint32_t yb_mask, uint32_t words[3]) {
  words[0] = __byte_perm(__byte_perm(c[0], c[1], 0x0073), __byte_perm(c[2], c[3], 0x0073), 0x5410);
  words[1] = __byte_perm(__byte_perm(c[4], c[5], 0x0073), __byte_perm(c[6], c[7], 0x0073), 0x5410);
  words[2] = __byte_perm(__byte_perm(c[8], c[9], 0x0073), __byte_perm(c[10], c[11], 0x0073), 0x5410);
  // distance certificates: the product drops the byte field and weighs the channel's distance by deltaF_max / deltaF_c
  const uint32_t wr = p.wmul[0], wg = p.wmul[1], wb = p.wmul[2], T = p.amb_t;
  const uint32_t d0 = min(min(c[0] * wr, c[1] * wg), c[2] * wb), d1 = min(min(c[3] * wr, c[4] * wg), c[5] * wb);
  const uint32_t d2 = min(min(c[6] * wr, c[7] * wg), c[8] * wb), d3 = min(min(c[9] * wr, c[10] * wg), c[11] * wb);
  uint32_t flags = 0;
  if (min(min(d0, d1), min(d2, d3)) <= T) {
    flags = (d0 <= T ? 1u : 0u) | (d1 <= T ? 2u : 0u) | (d2 <= T ? 4u : 0u) | (d3 <= T ? 8u : 0u);
  }
  if (fminf(ya, yb) < p.y_min)  // outside the certified domain (far below black): that pixel pair exactly
    flags |= (ya < p.y_min ? ya_mask : 0u) | (yb < p.y_min ? yb_mask : 0u);
  return flags;
}

// one four-pixel task of the cheap pass (RGB Bayer); returns the three output words and the mask of pixels to recompute
template <bool GF, bool AR>
__device__ __forceinline__ uint32_t cheap_task(const SpecParams &p, uint32_t g8_base, uint32_t stab_bias, const Window &w,
                                               uint32_t words[3]) {
  F2 a02, g02, o02, a13, g13, o13;
  demosaic_pairs<GF>(w, a02, g02, o02, a13, g13, o13);
  uint32_t s02[6], s13[6];
  const float y02 = chain_pair<false>(p, g8_base, stab_bias, AR ? a02 : o02, g02, AR ? o02 : a02, s02, nullptr);
  const float y13 = chain_pair<false>(p, g8_base, stab_bias, AR ? a13 : o13, g13, AR ? o13 : a13, s13, nullptr);
  // s[2c + h]: channel c of the pair's pixel h: px0 = s02[*][0], px1 = s13[*][0], px2 = s02[*][1], px3 = s13[*][1]
  const uint32_t c[12] = {s02[0], s02[2], s02[4], s13[0], s13[2], s13[4], s02[1], s02[3], s02[5], s13[1], s13[3], s13[5]};
  return pack_and_certify(p, c, y02, y13, 5u, 10u, words);
}

// the same task for any three-colour pattern: w = the 3 x 6 window (rows y-1 .. y+1, columns x0-1 .. x0+4), mm[j] = the
// tap masks of pixel j's pattern position (colours 0, 1 in .x low / high half, colour 2 in .y low half)
__device__ __forceinline__ uint32_t cheap_task_generic(const SpecParams &p, uint32_t g8_base, uint32_t stab_bias, const float w[3][6],
                                                       const uint2 mm[4], uint32_t words[3]) {
  float r[4], g[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float v[9];
#pragma unroll
    for (int k = 0; k < 3; k++) { v[k * 3] = w[k][j]; v[k * 3 + 1] = w[k][j + 1]; v[k * 3 + 2] = w[k][j + 2]; }
    r[j] = bin_mean_exact(mm[j].x & 0xffffu, v);
    g[j] = bin_mean_exact(mm[j].x >> 16, v);
    b[j] = bin_mean_exact(mm[j].y & 0xffffu, v);
  }
  uint32_t s01[6], s23[6];
  const float y01 = chain_pair<false>(p, g8_base, stab_bias, F2{r[0], r[1]}, F2{g[0], g[1]}, F2{b[0], b[1]}, s01, nullptr);
  const float y23 = chain_pair<false>(p, g8_base, stab_bias, F2{r[2], r[3]}, F2{g[2], g[3]}, F2{b[2], b[3]}, s23, nullptr);
  const uint32_t c[12] = {s01[0], s01[2], s01[4], s01[1], s01[3], s01[5], s23[0], s23[2], s23[4], s23[1], s23[3], s23[5]};
  return pack_and_certify(p, c, y01, y23, 3u, 12u, words);
}

__device__ __forceinline__ int atoms_add(uint32_t addr, int v) {  // plain shared-memory atomic (no warp aggregation)
  int old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}

// MODE 0..3: RGB Bayer, bit 1 = GF0 (even rows of the cropped frame start with green), bit 0 = AR0 (their other colour
// is red); odd rows are the opposite on both counts (green sits on one diagonal, red and blue on the other).
// MODE 4: any other three-colour pattern up to 12 x 12 (X-Trans): row-major tile, per-position tap masks.
// BATCH: the launch covers several frames (ipb_pipeline_output_8bit_batch); a single frame keeps the frame index and the
// per-frame output pointer out of its registers.
template <int NT, int MODE, bool BATCH>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 1)
k_spec8(const __grid_constant__ SpecParams p, const __grid_constant__ CfaDev cfa, const __grid_constant__ ColorParams P,
        const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemSpec &sm = *reinterpret_cast<SmemSpec *>(smem_raw);
  constexpr bool BAYER = MODE < 4, GF0 = (MODE & 2) != 0, AR0 = (MODE & 1) != 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = p.tiles_x * p.tiles_y * p.nframes;   // the frames of a batch are one sequence of tiles
  const uint32_t bar = smem_u32(&sm.mbar), bar_tab = smem_u32(&sm.mbar_tab), raw_addr = smem_u32(sm.raw);
  const uint8_t *pat = BAYER ? nullptr : sm.pat;   // the exact path's pattern kind

  // this CTA's tiles are blockIdx.x, + gridDim.x, ...: their positions by stepping (no division per tile)
  TilePos cur;
  {
    const int row = (int)blockIdx.x / p.tiles_x;
    cur.txi = (int)blockIdx.x - row * p.tiles_x;
    cur.f = BATCH ? row / p.tiles_y : 0;
    cur.tyi = row - cur.f * p.tiles_y;
  }
  const int step_y = (int)gridDim.x / p.tiles_x, step_x = (int)gridDim.x - step_y * p.tiles_x;
  auto advance = [&](TilePos &q) {
    q.txi += step_x; q.tyi += step_y;
    if (q.txi >= p.tiles_x) { q.txi -= p.tiles_x; q.tyi++; }
    if (BATCH)
      while (q.tyi >= p.tiles_y) { q.tyi -= p.tiles_y; q.f++; }
  };

  if (tid == 0) {
    sm.conv_ctr[0] = 0;
    sm.conv_ctr[1] = 0;
    sm.qn[0] = 0;
    sm.qn[1] = 0;
    mbar_init(bar, 1);
    mbar_init(bar_tab, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    issue_tile(p, &tmap, raw_addr, bar, cur);
    constexpr uint32_t kG8Bytes = kSpecG8Entries * 4u, kSTabBytes = kSpecSTabEntries * 8u;
    mbar_expect_tx(bar_tab, kG8Bytes + kSTabBytes);
    bulk_load(smem_u32(sm.g8a), p.g8a, kG8Bytes, bar_tab);
    bulk_load(smem_u32(sm.stab), p.stab, kSTabBytes, bar_tab);
  }
  const uint32_t g8_base = smem_u32(sm.g8a);
  if ((g8_base & 0x7fffu) != 0u) {  // cannot happen: ipb_ctx_create probed the shared window base (kSpecSmemBase)
    if (tid == 0 && p.stats) p.stats[4] = 1ull;
    return;
  }
  const uint32_t stab_bias = smem_u32(sm.stab) - p.bias58;  // (0x4B000000 << 3) mod 2^32: tf = 2^23 + key
  for (int i = tid; i < kSplRows; i += NT) {
    float e[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (i == 0) e[1] = P.sp.y_first;
    else if (i >= P.sp.n) e[1] = P.sp.y_last;
    else { e[0] = P.sp.x[i - 1]; e[1] = P.sp.y[i - 1]; e[2] = P.sp.c1[i - 1]; e[3] = P.sp.c2[i - 1]; e[4] = P.sp.c3[i - 1]; }
#pragma unroll
    for (int k = 0; k < 5; k++) sm.spl[i][k] = e[k];
  }
  const uint32_t phase = (uint32_t)cfa.pat[0] | ((uint32_t)cfa.pat[1] << 2) | ((uint32_t)cfa.pat[48] << 4) | ((uint32_t)cfa.pat[49] << 6);
  const uint32_t thr_base = smem_u32(sm.thr);
  if (!BAYER) {
    // demosaic.rs:77-90 for every position of the period: which of the nine 3x3 taps feed which colour (taps of the
    // centre's own colour other than the centre itself are discarded)
    for (int pos = tid; pos < p.pw * p.ph; pos += NT) {
      const int pr = pos / p.pw, pc = pos - pr * p.pw;
      const int pix = cfa.pat[pr * 48 + pc];
      uint32_t m[3] = {0u, 0u, 0u};
      int i = 0;
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++, i++) {
          const int oc = cfa.pat[((pr + 48 + dy) % 48) * 48 + (pc + 48 + dx) % 48];
          if ((oc != pix || (dx == 0 && dy == 0)) && oc < 3) m[oc] |= 1u << i;
        }
      sm.taps[pos] = make_uint2(m[0] | (m[1] << 16), m[2]);
      sm.pat[pos] = (uint8_t)pix;
    }
  }
  for (int i = tid; i < 260; i += NT) sm.thr[i] = i < 255 ? __ldg(p.thr8 + i) : __int_as_float(0x7f800000);
  for (int i = tid; i < (int)(sizeof(SpecParams) / 4); i += NT) reinterpret_cast<uint32_t *>(&sm.sp)[i] = reinterpret_cast<const uint32_t *>(&p)[i];
  for (int i = tid; i < (int)(sizeof(ColorParams) / 4); i += NT) reinterpret_cast<uint32_t *>(&sm.cp)[i] = reinterpret_cast<const uint32_t *>(&P)[i];

  // gofloat (gofloat.rs:127) of the staged raw box into tile buffer `buf`, exactly as the reference rounds it (the
  // cheap pass and the reference then start from identical samples).  Even / odd columns go to separate planes.
  // Warps pull chunks from a counter, so whichever warps finish their pixels first convert the next tile.
  auto convert_tile = [&](int buf, int ctr) {
    constexpr int kGroups = kTileElems / 8;          // 8 samples per thread per chunk
    const F2 rc = splat(p.range_rc), nrange = splat(-p.range), sub_a = splat(p.sub_a), sub_b = splat(p.sub_b);
    const bool two_subs = p.sub_b != 0.0f;
    for (;;) {
      int chunk = 0;
      if (lane == 0) chunk = atomicAdd(&sm.conv_ctr[ctr], 1);
      chunk = __shfl_sync(kFull, chunk, 0);
      if (chunk >= kChunks) break;
      const int gi = chunk * 32 + lane;
      if (gi < kGroups) {
        const uint4 pkd = *reinterpret_cast<const uint4 *>(sm.raw + gi * 8);
        const uint32_t w4[4] = {pkd.x, pkd.y, pkd.z, pkd.w};
        float ev[4], od[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          // {0x4B00 | lo16, 0x4B00 | hi16} as floats 2^23 + sample
          F2 t{__uint_as_float(__byte_perm(w4[k], 0x4B00u, 0x5410)), __uint_as_float(__byte_perm(w4[k], 0x4B00u, 0x5432))};
          // gofloat.rs:127, bit for bit: (v - black) is exact for an integral black level (one subtraction from
          // 2^23 + v), otherwise 2^23 comes off first; the division is the verified three-instruction form
          F2 num = a2(t, sub_a);
          if (two_subs) num = a2(num, sub_b);
          const F2 q = m2(num, rc);
          const F2 res = f2(f2(q, nrange, num), rc, q);
          ev[k] = fminf(res.x, 1.0f);
          od[k] = fminf(res.y, 1.0f);
        }
        const int r = gi / (kTileStride / 8), g = gi - r * (kTileStride / 8);
        float *row = &sm.plane[buf][r][0][0];
        if (BAYER) {
          *reinterpret_cast<float4 *>(row + 4 * g) = make_float4(ev[0], ev[1], ev[2], ev[3]);
          *reinterpret_cast<float4 *>(row + kPS + 4 * g) = make_float4(od[0], od[1], od[2], od[3]);
        } else {  // row-major
          *reinterpret_cast<float4 *>(row + 8 * g) = make_float4(ev[0], od[0], ev[1], od[1]);
          *reinterpret_cast<float4 *>(row + 8 * g + 4) = make_float4(ev[2], od[2], ev[3], od[3]);
        }
      }
    }
  };

  __syncthreads();
  mbar_wait(bar, 0);
  convert_tile(0, 0);
  __syncthreads();
  if (tid == 0 && (int)(blockIdx.x + gridDim.x) < ntiles) {
    TilePos nx = cur;
    advance(nx);
    issue_tile(p, &tmap, raw_addr, bar, nx);
  }
  mbar_wait(bar_tab, 0);

  // queue the uncertified pixels of a task (entry = row * kTW + column of its first pixel, inside the tile); the queue
  // holds a whole tile, so it cannot overflow
  auto push = [&](uint32_t qaddr, uint32_t flags, uint32_t entry) {
    int pos = atoms_add(qaddr, __popc(flags));
    while (flags) {
      const int j = __ffs(flags) - 1;
      flags &= flags - 1u;
      sm.queue[pos++] = (uint16_t)(entry + j);
    }
  };
  // whole words can be stored when every row starts on a 4-byte boundary
  const bool rows_aligned = ((reinterpret_cast<uintptr_t>(p.out) & 3) == 0) && ((p.width & 3) == 0) && ((p.frame_out_bytes & 3) == 0);

  int it = 0;
  int fix_warps = 0;                 // warps that recomputed pixels of the previous tile at the top of this iteration
  unsigned long long nfix = 0;       // thread 0: pixels recomputed by this CTA
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
    const TilePos tp = cur;
    advance(cur);
    const int ty0 = p.out_row0 + tp.tyi * kTH, tx0 = tp.txi * kTW;
    uint8_t *const out_f = BATCH ? p.out + (size_t)tp.f * (size_t)p.frame_out_bytes : p.out;   // this tile's frame of the batch
    const uint32_t tile_base = smem_u32(&sm.plane[it & 1][0][0][0]);
    const uint32_t qaddr = smem_u32(&sm.qn[it & 1]);
    // a tile is "inner" when all its pixels exist, are wanted, and have their nine taps inside the frame
    const bool inner = rows_aligned && ty0 >= 1 && ty0 + kTH <= p.height - 1 && ty0 + kTH <= p.out_row1 && tx0 >= 1 &&
                       tx0 + kTW <= p.width - 1;
    const uint32_t pix0 = (uint32_t)(ty0 - p.out_row0) * (uint32_t)p.width + (uint32_t)(tx0 + 4 * lane);

#pragma unroll 1
    for (int r = warp; r < kTH; r += NT / 32) {
      const int y = ty0 + r;
      uint32_t words[3], flags;
      if (BAYER) {
        // window: tile row r is frame row y-1; plane index of column x0 is 4 + 2 * lane
        Window w;
        const uint32_t a0 = tile_base + (uint32_t)(r * (2 * kPS) + 4 + 2 * lane) * 4u;
        constexpr uint32_t RS = 2 * kPS * 4, OP = kPS * 4;
        w.En = lds64f(a0); w.On = lds64f(a0 + OP);
        w.Ec = lds64f(a0 + RS); w.Oc = lds64f(a0 + RS + OP);
        w.Es = lds64f(a0 + 2 * RS); w.Os = lds64f(a0 + 2 * RS + OP);
        w.e2c = lds32f(a0 + RS + 8); w.omc = lds32f(a0 + RS + OP - 4);
        w.e2n = w.e2s = w.omn = w.oms = 0.0f;
        const bool gf = ((y & 1) != 0) != GF0;  // rows alternate
        if (gf) { w.e2n = lds32f(a0 + 8); w.e2s = lds32f(a0 + 2 * RS + 8); }
        else { w.omn = lds32f(a0 + OP - 4); w.oms = lds32f(a0 + 2 * RS + OP - 4); }
        if ((y & 1) == 0) flags = cheap_task<GF0, AR0>(p, g8_base, stab_bias, w, words);
        else flags = cheap_task<!GF0, !AR0>(p, g8_base, stab_bias, w, words);
      } else {
        // 3 x 6 window of the row-major tile: rows y-1 .. y+1, columns x0-1 .. x0+4 (tile column of x0 is 8 + 4 * lane)
        float w[3][6];
        const uint32_t a0 = tile_base + (uint32_t)(r * kTileStride + 7 + 4 * lane) * 4u;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const uint32_t a = a0 + (uint32_t)(k * kTileStride) * 4u;
          float4 mid;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(mid.x), "=f"(mid.y), "=f"(mid.z), "=f"(mid.w) : "r"(a + 4u));
          w[k][0] = lds32f(a);
          w[k][1] = mid.x; w[k][2] = mid.y; w[k][3] = mid.z; w[k][4] = mid.w;
          w[k][5] = lds32f(a + 20u);
        }
        // position inside the period: n % d as n - mulhi(n, ceil(2^32 / d)) * d
        const int xg = tx0 + 4 * lane;
        const int pr = y - (int)__umulhi((uint32_t)y, p.rcp_ph) * p.ph;
        int pc = xg - (int)__umulhi((uint32_t)xg, p.rcp_pw) * p.pw;
        uint2 mm[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          mm[j] = sm.taps[pr * p.pw + pc];
          pc = (pc + 1 == p.pw) ? 0 : pc + 1;
        }
        flags = cheap_task_generic(p, g8_base, stab_bias, w, mm, words);
      }
      const uint32_t pix = pix0 + (uint32_t)r * (uint32_t)p.width;
      if (inner) {
        uint32_t *o4 = reinterpret_cast<uint32_t *>(out_f + (size_t)pix * 3);
        o4[0] = words[0]; o4[1] = words[1]; o4[2] = words[2];
        if (flags && !(p.dbg & 2)) push(qaddr, flags, (uint32_t)(r * kTW + 4 * lane));
      } else {
        const int x0 = tx0 + 4 * lane;
        const bool live = y < p.out_row1 && x0 < p.width;
        if (live) {
          const int npx = min(4, p.width - x0);
          uint8_t *o = out_f + (size_t)pix * 3;
          if (npx == 4 && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
            uint32_t *o4 = reinterpret_cast<uint32_t *>(o);
            o4[0] = words[0]; o4[1] = words[1]; o4[2] = words[2];
          } else {
#pragma unroll
            for (int j = 0; j < 12; j++)
              if (j < npx * 3) o[j] = (uint8_t)(words[j >> 2] >> ((j & 3) * 8));
          }
          // frame border: the cheap demosaic assumed all nine taps (demosaic.rs:103-107 drops the missing ones)
          if (y < 1 || y > p.height - 2) flags = 0xfu;
          if (x0 < 1) flags |= 1u;
          if (x0 + 3 > p.width - 2) flags |= 0xfu & ~((1u << max(p.width - 1 - x0, 0)) - 1u);
          flags &= (1u << npx) - 1u;
          if (flags) push(qaddr, flags, (uint32_t)(r * kTW + 4 * lane));
        }
      }
    }

    const bool have_next = t + (int)gridDim.x < ntiles;
    // The conversion below overwrites the planes of the previous tile, which the warps that recompute that tile's
    // queued pixels are still reading: they arrive on barrier 1 when they are done, the other warps wait for them here
    // (seldom for long: recomputing takes about as long as one row of the cheap pass, and a warp has at least two).
    if (fix_warps > 0 && warp >= fix_warps) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    if (have_next) {
      mbar_wait(bar, (it + 1) & 1);
      convert_tile((it + 1) & 1, (it + 1) & 1);
    }
    if (tid == 0) {
      sm.conv_ctr[it & 1] = 0;
      sm.qn[(it + 1) & 1] = 0;   // the next tile's queue length; its last readers passed the second barrier of the previous iteration
    }
    __syncthreads();
    if (tid == 0 && t + 2 * (int)gridDim.x < ntiles) {
      TilePos nx = cur;   // already the next tile: one more step
      advance(nx);
      issue_tile(p, &tmap, raw_addr, bar, nx);
    }
    // Uncertified pixels of this tile are recomputed exactly from the tile's planes, densely: entry i goes to thread i, so
    // the first ceil(n / 32) warps do the work with all lanes busy while the others start the next tile.  The barrier
    // above ordered every push before these reads; the one below frees the queue for the next tile's pushes.
    const int qn = sm.qn[it & 1];
    const bool recompute = !(p.dbg & 1);
    if (qn > NT && recompute)   // more than one entry per thread (dark frames, the bound at its cap): all threads, right here
      for (int i = NT + tid; i < qn; i += NT)
        fixup_entry(sm.sp, sm.cp, sm.spl, phase, pat, sm.taps, g8_base, thr_base, tile_base, tx0, ty0, out_f, sm.queue[i]);
    const uint32_t mine = tid < qn ? (uint32_t)sm.queue[tid] : 0xffffffffu;
    __syncthreads();
    if (tid == 0) nfix += (unsigned long long)qn;
    fix_warps = min((qn + 31) >> 5, NT / 32);
    if (warp < fix_warps) {
      if (mine != 0xffffffffu && recompute)
        fixup_entry(sm.sp, sm.cp, sm.spl, phase, pat, sm.taps, g8_base, thr_base, tile_base, tx0, ty0, out_f, mine);
      if (have_next) asm volatile("bar.arrive 1, %0;" ::"n"(NT) : "memory");
    }
    if (!have_next) fix_warps = 0;
  }
  if (tid == 0 && p.stats && nfix) atomicAdd(p.stats, nfix);
}

// ---------------------------------------------------------------- speculative scaled kernel (BASELINE config 4)
// scaled_demosaic (scaling.rs:51-145) + the same cheap colour chain / certificate / exact recomputation for the 8-bit
// output of a down-scaled RGB Bayer frame.  The window phase is k_fused_scaled's (ipb_scaled.cuh: the reference's f32
// expressions, so cheap and exact chains start from identical demosaiced values); a thread takes two output pixels so
// that the chain runs on packed pairs.  Uncertified pixels go, with their demosaiced values, to a queue of the warp in
// shared memory; whenever it holds 32 the warp recomputes them exactly with all lanes busy (warp-synchronous, no CTA
// barrier), the rest after the last pixel.
constexpr int kScNT = 512;
constexpr int kWQCap = 96;   // 31 left over + up to 64 new entries per iteration

struct SmemScaledSpec : SmemFront {
  unsigned char fill[kG8Offset - (int)sizeof(SmemFront)];
  uint32_t g8a[kSpecG8Entries];
  float4 wq[kScNT / 32][kWQCap];                // {pixel index (bits), demosaiced r, g, b}
};
static_assert(offsetof(SmemScaledSpec, g8a) == kG8Offset, "gamma table must sit on a 32 KB boundary of the shared window");

__device__ __noinline__ void fixup_scaled(const SpecParams &p, const ColorParams &P, const float (*spl)[8], uint32_t g8_base,
                                          uint32_t thr_base, float4 e) {
  float v[3];
  exact_chain(p, P, spl, e.y, e.z, e.w, v);
  uint8_t *o = p.out + (size_t)__float_as_uint(e.x) * 3;
  o[0] = (uint8_t)gamma8_exact(g8_base, thr_base, v[0]);
  o[1] = (uint8_t)gamma8_exact(g8_base, thr_base, v[1]);
  o[2] = (uint8_t)gamma8_exact(g8_base, thr_base, v[2]);
}

__global__ void __launch_bounds__(kScNT, 2)
k_spec8_scaled(const __grid_constant__ SpecParams p, const __grid_constant__ ScaledParams g, const __grid_constant__ CfaDev cfa,
               const __grid_constant__ ColorParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemScaledSpec &sm = *reinterpret_cast<SmemScaledSpec *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar_tab = smem_u32(&sm.mbar_tab);
  if (tid == 0) {
    mbar_init(bar_tab, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    constexpr uint32_t kG8Bytes = kSpecG8Entries * 4u, kSTabBytes = kSpecSTabEntries * 8u;
    mbar_expect_tx(bar_tab, kG8Bytes + kSTabBytes);
    bulk_load(smem_u32(sm.g8a), p.g8a, kG8Bytes, bar_tab);
    bulk_load(smem_u32(sm.stab), p.stab, kSTabBytes, bar_tab);
  }
  const uint32_t g8_base = smem_u32(sm.g8a);
  if ((g8_base & 0x7fffu) != 0u) {  // cannot happen: ipb_ctx_create probed the shared window base (kSpecSmemBase)
    if (tid == 0 && p.stats) p.stats[4] = 1ull;
    return;
  }
  const uint32_t stab_bias = smem_u32(sm.stab) - p.bias58, thr_base = smem_u32(sm.thr);
  for (int i = tid; i < kSplRows; i += kScNT) {
    float e[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (i == 0) e[1] = P.sp.y_first;
    else if (i >= P.sp.n) e[1] = P.sp.y_last;
    else { e[0] = P.sp.x[i - 1]; e[1] = P.sp.y[i - 1]; e[2] = P.sp.c1[i - 1]; e[3] = P.sp.c2[i - 1]; e[4] = P.sp.c3[i - 1]; }
#pragma unroll
    for (int k = 0; k < 5; k++) sm.spl[i][k] = e[k];
  }
  for (int i = tid; i < 260; i += kScNT) sm.thr[i] = i < 255 ? __ldg(p.thr8 + i) : __int_as_float(0x7f800000);
  for (int i = tid; i < (int)(sizeof(SpecParams) / 4); i += kScNT) reinterpret_cast<uint32_t *>(&sm.sp)[i] = reinterpret_cast<const uint32_t *>(&p)[i];
  for (int i = tid; i < (int)(sizeof(ColorParams) / 4); i += kScNT) reinterpret_cast<uint32_t *>(&sm.cp)[i] = reinterpret_cast<const uint32_t *>(&P)[i];
  __syncthreads();

  float4 *wq = sm.wq[warp];
  int qn = 0;                        // warp-uniform: entries in this warp's queue
  unsigned long long nfix = 0;
  bool tables_in = false;
  const bool recompute = !(p.dbg & 1);
  // a thread takes the two pixels of one column in rows 2k, 2k + 1 of the launch: consecutive lanes on consecutive
  // columns (their window loads touch the fewest cache lines), and the column part of the window computed once
  const long long npix = (long long)(g.out_row1 - g.out_row0) * g.nwidth;
  const int nrows = g.out_row1 - g.out_row0;
  const long long npairs = (long long)((nrows + 1) / 2) * g.nwidth, npairs_pad = (npairs + 31) / 32 * 32;  // whole warps stay in the loop
  for (long long pi = (long long)blockIdx.x * kScNT + tid; pi < npairs_pad; pi += (long long)gridDim.x * kScNT) {
    const bool live = pi < npairs;
    const long long pj = live ? pi : npairs - 1;
    const int rp = (int)(pj / g.nwidth), col = (int)(pj - (long long)rp * g.nwidth);
    const bool have1 = live && 2 * rp + 1 < nrows;
    const long long id0 = (long long)(2 * rp) * g.nwidth + col, id1 = have1 ? id0 + g.nwidth : id0;
    float pa[3], pb[3];
    scaled_pair_bayer(g, cfa, g.out_row0 + 2 * rp, col, 2 * rp + 1 < nrows, pa, pb);
    if (!tables_in) {
      mbar_wait(bar_tab, 0);
      tables_in = true;
    }
    // a weighted mean of samples <= 1 can round an ulp above 1, where the reference clips green (mul[1] == 1); the cheap
    // chain assumes green <= 1
    uint32_t s[6];
    const float ymin = chain_pair<false>(p, g8_base, stab_bias, F2{pa[0], pb[0]}, F2{fminf(pa[1], 1.0f), fminf(pb[1], 1.0f)},
                                         F2{pa[2], pb[2]}, s, nullptr);
    const uint32_t wr = p.wmul[0], wg = p.wmul[1], wb = p.wmul[2], T = p.amb_t;
    const uint32_t d0 = min(min(s[0] * wr, s[2] * wg), s[4] * wb), d1 = min(min(s[1] * wr, s[3] * wg), s[5] * wb);
    const bool dark = ymin < p.y_min;   // outside the certified domain: both pixels exactly
    const bool f0 = live && (d0 <= T || dark) && !(p.dbg & 2), f1 = have1 && (d1 <= T || dark) && !(p.dbg & 2);
    if (live) {
      uint8_t *o = p.out + (size_t)id0 * 3;
      o[0] = (uint8_t)(s[0] >> 24); o[1] = (uint8_t)(s[2] >> 24); o[2] = (uint8_t)(s[4] >> 24);
      if (have1) {
        uint8_t *o1 = p.out + (size_t)id1 * 3;
        o1[0] = (uint8_t)(s[1] >> 24); o1[1] = (uint8_t)(s[3] >> 24); o1[2] = (uint8_t)(s[5] >> 24);
      }
    }
    // queue the uncertified pixels with their demosaiced values (ballot compaction: qn stays warp-uniform)
    const uint32_t m0 = __ballot_sync(kFull, f0), m1 = __ballot_sync(kFull, f1), lt = (1u << lane) - 1u;
    if (f0) wq[qn + __popc(m0 & lt)] = make_float4(__uint_as_float((uint32_t)id0), pa[0], pa[1], pa[2]);
    qn += __popc(m0);
    if (f1) wq[qn + __popc(m1 & lt)] = make_float4(__uint_as_float((uint32_t)id1), pb[0], pb[1], pb[2]);
    qn += __popc(m1);
    __syncwarp();   // orders the cheap stores and the queue writes before the recomputation by other lanes
    while (qn >= 32) {
      qn -= 32;
      const float4 e = wq[qn + lane];
      if (recompute) fixup_scaled(sm.sp, sm.cp, sm.spl, g8_base, thr_base, e);
      nfix += 32;
      __syncwarp();
    }
  }
  if (qn > 0) {
    if (lane < qn && recompute) fixup_scaled(sm.sp, sm.cp, sm.spl, g8_base, thr_base, wq[lane]);
    nfix += (unsigned long long)qn;
  }
  if (lane == 0 && p.stats && nfix) atomicAdd(p.stats, nfix);
}

// ---------------------------------------------------------------- probe: cheap vs exact linear values
// One thread per interior four-pixel task straight from global memory; stats[1] = max |cheap - exact| over the
// clamped linear channel values (as float bits), stats[2] = number of values compared, stats[3] = sum of |diff| * 2^40.
__global__ void k_spec_probe(const __grid_constant__ SpecParams p, const __grid_constant__ CfaDev cfa,
                             const __grid_constant__ ColorParams P) {
  extern __shared__ __align__(16) unsigned char probe_smem[];
  uint32_t *g8a = reinterpret_cast<uint32_t *>(probe_smem + kG8Offset);
  float2 *stab = reinterpret_cast<float2 *>(probe_smem);
  for (int i = threadIdx.x; i < kSpecG8Entries; i += blockDim.x) g8a[i] = p.g8a[i];
  for (int i = threadIdx.x; i < kSpecSTabEntries; i += blockDim.x) stab[i] = p.stab[i];
  float (*spl)[8] = reinterpret_cast<float (*)[8]>(probe_smem + kSpecSTabEntries * 8);
  for (int i = threadIdx.x; i < kSplRows; i += blockDim.x) {
    float e[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (i == 0) e[1] = P.sp.y_first;
    else if (i >= P.sp.n) e[1] = P.sp.y_last;
    else { e[0] = P.sp.x[i - 1]; e[1] = P.sp.y[i - 1]; e[2] = P.sp.c1[i - 1]; e[3] = P.sp.c2[i - 1]; e[4] = P.sp.c3[i - 1]; }
    for (int k = 0; k < 5; k++) spl[i][k] = e[k];
  }
  const uint32_t phase = (uint32_t)cfa.pat[0] | ((uint32_t)cfa.pat[1] << 2) | ((uint32_t)cfa.pat[48] << 4) | ((uint32_t)cfa.pat[49] << 6);
  __syncthreads();
  const uint32_t g8_base = smem_u32(g8a), stab_bias = smem_u32(stab) - p.bias58;
  if ((g8_base & 0x7fffu) != 0u) return;
  const int tasks_x = (p.width - 2) / 4;  // tasks start at x0 = 4, 8, ... (aligned like the kernel's), interior only
  const long long ntasks = (long long)tasks_x * (p.height - 2);
  float worst = 0.0f;
  unsigned long long cnt = 0, sum = 0;
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < ntasks; id += (long long)gridDim.x * blockDim.x) {
    const int y = 1 + (int)(id / tasks_x), x0 = 4 * (1 + (int)(id % tasks_x));
    if (x0 + 4 > p.width - 1) continue;
    auto ld = [&](int yy, int xx) {
      const float v = (float)__ldg(p.raw + (long long)(yy + p.crop_y - p.src_row0) * p.raw_pitch + p.crop_x + xx);
      return golevel(v, p.black, p.range, p.range_rc, p.exact_rc);
    };
    Window w;
    w.En = F2{ld(y - 1, x0), ld(y - 1, x0 + 2)}; w.On = F2{ld(y - 1, x0 + 1), ld(y - 1, x0 + 3)};
    w.Ec = F2{ld(y, x0), ld(y, x0 + 2)}; w.Oc = F2{ld(y, x0 + 1), ld(y, x0 + 3)};
    w.Es = F2{ld(y + 1, x0), ld(y + 1, x0 + 2)}; w.Os = F2{ld(y + 1, x0 + 1), ld(y + 1, x0 + 3)};
    w.e2n = ld(y - 1, x0 + 4); w.e2c = ld(y, x0 + 4); w.e2s = ld(y + 1, x0 + 4);
    w.omn = ld(y - 1, x0 - 1); w.omc = ld(y, x0 - 1); w.oms = ld(y + 1, x0 - 1);
    const bool odd = (y & 1) != 0;
    const bool gf = cfa.pat[(odd ? 48 : 0)] == 1;
    const bool ar = (gf ? cfa.pat[(odd ? 48 : 0) + 1] : cfa.pat[odd ? 48 : 0]) == 0;
    F2 a02, g02, o02, a13, g13, o13;
    if (gf) demosaic_pairs<true>(w, a02, g02, o02, a13, g13, o13);
    else demosaic_pairs<false>(w, a02, g02, o02, a13, g13, o13);
    uint32_t s02[6], s13[6];
    float l02[6], l13[6];
    const float y02 = chain_pair<true>(p, g8_base, stab_bias, ar ? a02 : o02, g02, ar ? o02 : a02, s02, l02);
    const float y13 = chain_pair<true>(p, g8_base, stab_bias, ar ? a13 : o13, g13, ar ? o13 : a13, s13, l13);
    if (fminf(y02, y13) < p.y_min) continue;  // outside the certified domain: the kernel recomputes these
    for (int j = 0; j < 4; j++) {
      float ex[3];
      float tp[9];
      taps_from_frame(p, x0 + j, y, tp);
      float er, eg, eb;
      exact_rgb_bayer(p, phase, x0 + j, y, tp, er, eg, eb);
      exact_chain(p, P, spl, er, eg, eb, ex);
      for (int c = 0; c < 3; c++) ex[c] = fminf(fmaxf(ex[c], 0.0f), 1.0f);  // gamma.rs:21 clamps before the table
      const float *l = (j & 1) ? l13 : l02;
      const int h = j >> 1;
      for (int c = 0; c < 3; c++) {
        const float d = fabsf(l[2 * c + h] - ex[c]);
        worst = fmaxf(worst, d);
        sum += (unsigned long long)(d * 1099511627776.0f);
        cnt++;
      }
    }
  }
  atomicMax(reinterpret_cast<unsigned int *>(p.stats) + 2, __float_as_uint(worst));  // stats[1] low word
  atomicAdd(p.stats + 2, cnt);
  atomicAdd(p.stats + 3, sum);
}

// Self-test run once per context: out[0] = shared-window address of dynamic shared memory (the gamma table's
// one-instruction addressing assumes kSpecSmemBase), out[1] = max relative error, as float bits, of the XU-pipe cube
// root ex2(lg2(v)/3) against the correctly rounded one over EVERY float in [2^-8, 4] — the measured constant the
// error bound of the cheap pass uses for its Lab transfer function.
__global__ void k_spec_selftest(unsigned int *out) {
  extern __shared__ __align__(128) unsigned char st_smem[];
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = smem_u32(st_smem);
  float worst = 0.0f;
  const uint32_t lo = 0x3b800000u, hi = 0x40800000u;  // 2^-8 .. 4.0
  for (uint32_t b = lo + blockIdx.x * blockDim.x + threadIdx.x; b <= hi; b += gridDim.x * blockDim.x) {
    const float v = __uint_as_float(b);
    const float got = ex2a(lg2a(v) * (1.0f / 3.0f));
    const double want = cbrt((double)v);
    worst = fmaxf(worst, (float)(fabs((double)got - want) / want));
  }
  atomicMax(out + 1, __float_as_uint(worst));
}

thread_local const char *g_spec_err = "";

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
bool make_raw_tmap(CUtensorMap *map, const uint16_t *raw, size_t pitch_elems, size_t rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(raw) & 15) || ((pitch_elems * sizeof(uint16_t)) & 15) || rows == 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)pitch_elems, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)(pitch_elems * sizeof(uint16_t))};
  const cuuint32_t box[2] = {(cuuint32_t)kTileStride, (cuuint32_t)kTileRows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t *>(raw), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NT, int MODE, bool BATCH>
cudaError_t launch_variant_b(cudaStream_t s, const SpecParams &p, const CfaDev &cfa, const ColorParams &P,
                             const CUtensorMap &tmap, int ntiles, int sm_count) {
  const size_t smem = sizeof(SmemSpec);
  const int ctas = sm_count * (NT == 512 ? 2 : 1);
  const int grid = ntiles < ctas ? ntiles : ctas;
  cudaError_t e = cudaFuncSetAttribute(k_spec8<NT, MODE, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_spec8<NT, MODE, BATCH><<<grid, NT, smem, s>>>(p, cfa, P, tmap);
  return cudaGetLastError();
}
template <int NT, int MODE>
cudaError_t launch_variant(cudaStream_t s, const SpecParams &p, const CfaDev &cfa, const ColorParams &P,
                           const CUtensorMap &tmap, int ntiles, int sm_count) {
  // batches only at the default CTA size (the 1024-thread variant is a measurement aid)
  if (p.nframes > 1 && NT == 512) return launch_variant_b<512, MODE, true>(s, p, cfa, P, tmap, ntiles, sm_count);
  if (p.nframes > 1) return cudaErrorInvalidValue;
  return launch_variant_b<NT, MODE, false>(s, p, cfa, P, tmap, ntiles, sm_count);
}

// RGB Bayer: 2 x 2, green on one diagonal, red and blue on the other
bool is_rgb_bayer(const CfaDev &cfa) {
  if (cfa.width != 2 || cfa.height != 2) return false;
  const int c0 = cfa.pat[0], c1 = cfa.pat[1], c2_ = cfa.pat[48], c3 = cfa.pat[49];
  return (c1 == 1 && c2_ == 1 && ((c0 == 0 && c3 == 2) || (c0 == 2 && c3 == 0))) ||
         (c0 == 1 && c3 == 1 && ((c1 == 0 && c2_ == 2) || (c1 == 2 && c2_ == 0)));
}
// any other pattern of colours 0..2 with a period the tap-mask table holds (X-Trans 6 x 6, 2 x 8, 12 x 12 ...)
bool is_rgb_generic(const CfaDev &cfa) {
  if (cfa.width < 2 || cfa.height < 2 || cfa.width > 12 || cfa.height > 12) return false;
  if (48 % cfa.width != 0 || 48 % cfa.height != 0) return false;  // the 48 x 48 table must tile without a seam
  for (int r = 0; r < cfa.height; r++)
    for (int c = 0; c < cfa.width; c++)
      if (cfa.pat[r * 48 + c] > 2) return false;
  return true;
}

}  // namespace

const char *spec_last_error() { return g_spec_err; }

bool spec_supported(const FusedArgs &a, const CfaDev &cfa, const ColorParams &P) {
  if (a.out_kind != kOutU8 || P.linear || P.use_e) return false;
  if (!is_rgb_bayer(cfa) && !is_rgb_generic(cfa)) return false;
  if (!a.use_tma || (a.crop_x % 8) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a.raw) & 15) || ((a.raw_pitch * sizeof(uint16_t)) & 15)) return false;
  if (a.width < 8 || a.height < 4) return false;
  if (a.width >= 65536 || a.out_row1 - a.out_row0 >= 65536) return false;  // queue entries are row << 16 | column
  if (a.height >= (1u << 24)) return false;  // n % period by multiplication (k_spec8, generic patterns)
  if (a.batch_n > 1) {   // tile numbers and tensor-map rows of the whole batch stay inside 31 bits
    const unsigned long long tiles = (unsigned long long)((a.width + kTW - 1) / kTW) * ((a.out_row1 - a.out_row0 + kTH - 1) / kTH);
    if (tiles * a.batch_n >= (1ull << 30) || (unsigned long long)a.batch_src_rows * a.batch_n >= (1ull << 30)) return false;
    if (a.batch_src_rows < a.src_rows) return false;   // frames must not overlap
  }
  return true;
}

cudaError_t launch_fused_spec8(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                               const SpecTables &T, int sm_count, int threads) {
  if (a.out_row1 <= a.out_row0 || a.width == 0) return cudaSuccess;
  SpecParams p = T.consts;
  p.raw = a.raw; p.raw_pitch = (long long)a.raw_pitch;
  p.src_row0 = (int)a.src_row0; p.src_rows = (int)a.src_rows;
  p.crop_x = (int)a.crop_x; p.crop_y = (int)a.crop_y;
  p.width = (int)a.width; p.height = (int)a.height;
  p.out_row0 = (int)a.out_row0; p.out_row1 = (int)a.out_row1;
  p.out = (uint8_t *)a.out;
  p.black = a.black; p.range = a.range; p.range_rc = a.range_rc; p.exact_rc = a.exact_rc;
  {
    const bool integral = a.black >= 0.0f && a.black < 4194304.0f && a.black == floorf(a.black);
    p.sub_a = integral ? -(8388608.0f + a.black) : -8388608.0f;
    p.sub_b = integral ? 0.0f : -a.black;
  }
  p.bias58 = 0x58000000u;
  {
    static const int dbg = getenv("IPB_SPEC_DBG") ? atoi(getenv("IPB_SPEC_DBG")) : 0;
    p.dbg = dbg;
  }
  p.lut_lab = a.lut_lab; p.lut_gamma = a.lut_gamma; p.cbrt_tab = a.cbrt_tab;
  p.g8a = T.g8a; p.stab = T.stab; p.thr8 = T.thr8; p.stats = T.stats;
  p.tiles_x = (p.width + kTW - 1) / kTW;
  p.tiles_y = (p.out_row1 - p.out_row0 + kTH - 1) / kTH;
  p.nframes = a.batch_n > 1 ? (int)a.batch_n : 1;
  p.frame_src_rows = p.nframes > 1 ? (int)a.batch_src_rows : 0;
  p.frame_out_bytes = p.nframes > 1 ? (long long)a.batch_out_bytes : 0;
  // one tensor map over the rows of all frames: a box that reaches into the neighbouring frame (or past the last one:
  // zero fill) only fetches taps that the frame-border logic never uses
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (!make_raw_tmap(&tmap, a.raw, a.raw_pitch, (size_t)(p.nframes - 1) * (size_t)p.frame_src_rows + a.src_rows)) {
    g_spec_err = "spec8: tensor map";
    return cudaErrorInvalidValue;
  }
  const int ntiles = p.tiles_x * p.tiles_y * p.nframes;
  p.pw = cfa.width; p.ph = cfa.height;
  p.rcp_pw = (uint32_t)(0x100000000ull / (unsigned long long)cfa.width) + 1u;
  p.rcp_ph = (uint32_t)(0x100000000ull / (unsigned long long)cfa.height) + 1u;
  if (p.nframes > 1) threads = 512;
  if (!is_rgb_bayer(cfa)) {
    if (threads == 1024) return launch_variant<1024, 4>(s, p, cfa, P, tmap, ntiles, sm_count);
    return launch_variant<512, 4>(s, p, cfa, P, tmap, ntiles, sm_count);
  }
  const bool gf0 = cfa.pat[0] == 1;
  const bool ar0 = (gf0 ? cfa.pat[1] : cfa.pat[0]) == 0;
  const int variant = (threads == 1024 ? 4 : 0) | (gf0 ? 2 : 0) | (ar0 ? 1 : 0);
  switch (variant) {
    case 0: return launch_variant<512, 0>(s, p, cfa, P, tmap, ntiles, sm_count);
    case 1: return launch_variant<512, 1>(s, p, cfa, P, tmap, ntiles, sm_count);
    case 2: return launch_variant<512, 2>(s, p, cfa, P, tmap, ntiles, sm_count);
    case 3: return launch_variant<512, 3>(s, p, cfa, P, tmap, ntiles, sm_count);
    case 4: return launch_variant<1024, 0>(s, p, cfa, P, tmap, ntiles, sm_count);
    case 5: return launch_variant<1024, 1>(s, p, cfa, P, tmap, ntiles, sm_count);
    case 6: return launch_variant<1024, 2>(s, p, cfa, P, tmap, ntiles, sm_count);
    default: return launch_variant<1024, 3>(s, p, cfa, P, tmap, ntiles, sm_count);
  }
}

bool spec_scaled_supported(const FusedArgs &a, const CfaDev &cfa, const ColorParams &P) {
  if (a.out_kind != kOutU8 || P.linear || P.use_e || !a.exact_rc) return false;
  if (!is_rgb_bayer(cfa)) return false;
  if (a.out_width < 2 || a.out_height < 2 || a.width < 2 || a.height < 2) return false;
  // windows of at most kMaxCols columns (scaling.rs:84-87: floor(skip * (col + 1)) - floor(skip * col) + 1 <= ceil(skip) + 1)
  const float skip_x = (float)((long)a.width - 1) / (float)(a.out_width - 1);
  if (!(skip_x >= 1.0f) || ceilf(skip_x) + 1.0f > (float)kMaxCols) return false;
  if ((unsigned long long)a.out_width * (unsigned long long)(a.out_row1 - a.out_row0) >= (1ull << 31)) return false;
  return true;
}

cudaError_t launch_scaled_spec8(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                                const SpecTables &T, int sm_count) {
  if (a.out_row1 <= a.out_row0 || a.out_width == 0) return cudaSuccess;
  SpecParams p = T.consts;
  p.out = (uint8_t *)a.out;
  p.black = a.black; p.range = a.range; p.range_rc = a.range_rc; p.exact_rc = a.exact_rc;
  p.bias58 = 0x58000000u;
  {
    static const int dbg = getenv("IPB_SPEC_DBG") ? atoi(getenv("IPB_SPEC_DBG")) : 0;
    p.dbg = dbg;
  }
  p.lut_lab = a.lut_lab; p.lut_gamma = a.lut_gamma; p.cbrt_tab = a.cbrt_tab;
  p.g8a = T.g8a; p.stab = T.stab; p.thr8 = T.thr8; p.stats = T.stats;
  ScaledParams g;
  fill_scaled_params(a, cfa, &g);
  const long long npairs = (long long)((g.out_row1 - g.out_row0 + 1) / 2) * g.nwidth;
  const long long blocks = (npairs + kScNT - 1) / kScNT;
  const int grid = (int)(blocks < 2ll * sm_count ? blocks : 2ll * sm_count);
  const size_t smem = sizeof(SmemScaledSpec);
  cudaError_t e = cudaFuncSetAttribute(k_spec8_scaled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_spec8_scaled<<<grid, kScNT, smem, s>>>(p, g, cfa, P);
  return cudaGetLastError();
}

cudaError_t launch_spec_selftest(cudaStream_t s, unsigned int *out2) {
  const size_t smem = sizeof(SmemSpec);
  cudaError_t e = cudaFuncSetAttribute(k_spec_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_spec_selftest<<<592, 512, smem, s>>>(out2);
  return cudaGetLastError();
}

cudaError_t launch_spec_probe(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                              const SpecTables &T, int sm_count) {
  // Bayer frames only: the cheap demosaic of every pattern is exact, so what the probe measures — the colour chain from
  // the demosaiced values on — does not depend on the pattern
  if (!is_rgb_bayer(cfa)) {
    g_spec_err = "spec probe: RGB Bayer frames only";
    return cudaErrorInvalidValue;
  }
  SpecParams p = T.consts;
  p.raw = a.raw; p.raw_pitch = (long long)a.raw_pitch;
  p.src_row0 = (int)a.src_row0; p.src_rows = (int)a.src_rows;
  p.crop_x = (int)a.crop_x; p.crop_y = (int)a.crop_y;
  p.width = (int)a.width; p.height = (int)a.height;
  p.out_row0 = 0; p.out_row1 = (int)a.height;
  p.out = nullptr;
  p.nframes = 1; p.frame_src_rows = 0; p.frame_out_bytes = 0;
  p.black = a.black; p.range = a.range; p.range_rc = a.range_rc; p.exact_rc = a.exact_rc;
  {
    const bool integral = a.black >= 0.0f && a.black < 4194304.0f && a.black == floorf(a.black);
    p.sub_a = integral ? -(8388608.0f + a.black) : -8388608.0f;
    p.sub_b = integral ? 0.0f : -a.black;
  }
  p.bias58 = 0x58000000u;
  p.lut_lab = a.lut_lab; p.lut_gamma = a.lut_gamma; p.cbrt_tab = a.cbrt_tab;
  p.g8a = T.g8a; p.stab = T.stab; p.thr8 = T.thr8; p.stats = T.stats;
  const size_t smem = kG8Offset + kSpecG8Entries * 4;
  cudaError_t e = cudaFuncSetAttribute(k_spec_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_spec_probe<<<sm_count * 2, 256, smem, s>>>(p, cfa, P);
  return cudaGetLastError();
}

}  // namespace ipb
