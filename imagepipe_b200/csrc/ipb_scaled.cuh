// ipb_scaled.cuh — the window phase of scaled_demosaic (scaling.rs:51-145) shared by k_fused_scaled (ipb_fused.cu) and
// the speculative scaled kernel (ipb_spec.cu): parameters, coordinate arithmetic and the tap loops, in exactly the
// reference's f32 expression order.
#pragma once
#include "ipb_internal.h"

namespace ipb {

// ---------------------------------------------------------------------------------------- scaled demosaic

struct ScaledParams {
  const uint16_t *raw;
  long long raw_pitch;
  int src_row0, src_rows;
  int crop_x, crop_y;
  int width, height;            // cropped source frame
  int nwidth, nheight;          // output frame
  int out_row0, out_row1;
  void *out;
  float black, range, range_rc;
  int exact_rc;
  const float2 *lut_lab, *lut_gamma;
  float skip_x, skip_y;         // skip_x_x, skip_y_y of scaling.rs:69-72 (skip_x_y == skip_y_x == 0 here)
  float skip_x_rc, skip_y_rc;   // their reciprocals
  int skip_rc_exact;            // the reciprocal form of delta / skip equals IEEE division for every tap of this frame
                                // (checked on the host over all of them: scaled_skip_rc_exact)
  float sub_a, sub_b;           // level mapping: (2^23 + v) + sub_a [+ sub_b] == (float)v - black exactly (see launch)
  int bayer;                    // 2x2 pattern with green on one diagonal and red / blue on the other
};

constexpr int kPatStride = 56;   // pattern row: 48 columns + the first 8 again, so that x % 48 + k needs no wrap
constexpr int kMaxCols = 8;      // widest window the register-resident fast path handles (scale < 7)


__device__ __forceinline__ int f2i_sat(float f) {  // Rust `f as usize` for the values met here (>= 0, < 2^31)
  return (int)min(__float2uint_rz(f), 0x7fffffffu);
}

// sums[c] += vf; counts[c] += f when cond, as two predicated scalar adds (a predicated packed add is lowered to an
// unpredicated FFMA2 plus two selects, which costs more issue slots than this)
__device__ __forceinline__ void add_if(F2 &acc, float vf, float f, bool cond) {
  asm("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q add.rn.f32 %0, %0, %3; @q add.rn.f32 %1, %1, %4; }"
      : "+f"(acc.x), "+f"(acc.y) : "r"((int)cond), "f"(vf), "f"(f));
}

// u16 -> f32 without the conversion unit: 0x4B000000 | v is the float 2^23 + v, and subtracting 2^23 is exact
__device__ __forceinline__ float u16_to_float(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }

// The taps of one output pixel whose window is at most NX columns wide (scaling.rs:91-118, CFA mode).  Columns k >= nx
// of a narrower window (frame edge, or a lane whose window is narrower than its warp's) get weight 0 and re-read the
// window's last column: they add (+-0, 0) to the accumulators, which changes neither a sum nor a count (neither is
// ever -0.0: both start at +0.0).
template <int NX, bool UNIFORM, int NC>
__device__ __forceinline__ void window_taps(const ScaledParams &p, const uint8_t *pat, float one, int from_x, int nx,
                                            int from_y, int to_y, float center_x, float center_y, F2 acc[4]) {
  const float black = p.black, range = p.range, rc = p.range_rc, skip_x = p.skip_x, skip_y = p.skip_y;
  float ax[NX];  // 1.0 - delta_x*delta_x of window column k
#pragma unroll
  for (int k = 0; k < NX; k++) {
    const float delta_x = __fdiv_rn((float)(from_x + k) - center_x, skip_x);
    ax[k] = 1.0f - (delta_x * delta_x);
  }
  int ym = from_y % 48;
  const uint8_t *pcol = pat + from_x % 48;
  const uint16_t *rowp = p.raw + (long long)(from_y + p.crop_y - p.src_row0) * p.raw_pitch + p.crop_x + from_x;
  for (int y = from_y; y <= to_y; y++, rowp += p.raw_pitch) {
    const float delta_y = __fdiv_rn((float)y - center_y, skip_y);
    const float dy2 = delta_y * delta_y;
    const uint8_t *prow = pcol + ym * kPatStride;
    ym = ym == 47 ? 0 : ym + 1;
#pragma unroll
    for (int k = 0; k < NX; k++) {
      const bool valid = UNIFORM || k < nx;
      float factor = ax[k] - dy2;
      factor = factor < 0.0f ? 0.0f : factor;
      if (!UNIFORM) factor = valid ? factor : 0.0f;
      const int kk = UNIFORM ? k : min(k, nx - 1);
      const int c = prow[kk];
      const float v = fminf(div_rc(u16_to_float(__ldg(rowp + kk)) - black, range, rc), 1.0f);  // gofloat.rs:127
      const float vf = v * factor;
#pragma unroll
      for (int j = 0; j < NC; j++) add_if(acc[j], vf, factor, c == j);
    }
  }
}

// One window row of an RGB Bayer frame.  RP = parity of the row relative to the window's first row.  In a Bayer mosaic
// green sits on one diagonal, so whether window column k of this row is green depends only on (RP + k) & 1 and on
// one per-lane bit — is the window's top-left sample green (g00) — and the row's other colour is the same for the
// whole row.  Green taps go to `g`, the others to `xacc` (the caller keeps one per relative row parity and maps the
// two to red / blue at the end): no colour look-up, no comparisons, four predicated adds per tap.  Each colour still
// receives its taps in raster order, like the reference's sums[c] / counts[c].
template <int NX, bool UNIFORM, int RP>
__device__ __forceinline__ void bayer_row(const uint32_t smp[NX], const float ax[NX], float dy2, int nx, bool g00, float sub_a,
                                          float sub_b, float range, float rc, F2 &g, F2 &xacc) {
#pragma unroll
  for (int k = 0; k < NX; k++) {
    // scaling.rs:106-107; fmaxf == `if factor < 0.0 {0.0} else {factor}` here: the factor is never NaN (finite geometry)
    // and a zero of either sign adds nothing to sums that start at +0.0
    float factor = fmaxf(ax[k] - dy2, 0.0f);
    if (!UNIFORM) factor = k < nx ? factor : 0.0f;
    // gofloat.rs:127: (v - black) is one exact subtraction from 2^23 + v for an integral black level, two otherwise
    float num = __uint_as_float(0x4B000000u | smp[k]) + sub_a;
    if (sub_b != 0.0f) num = num + sub_b;
    const float v = fminf(div_rc(num, range, rc), 1.0f);
    const float vf = v * factor;
    // (measured: loops specialised for a warp-uniform phase, with plain instead of predicated adds, run slower)
    const bool green = ((RP + k) & 1) ? !g00 : g00;
    add_if(g, vf, factor, green);
    add_if(xacc, vf, factor, !green);
  }
}

// SRC: delta / skip in the verified reciprocal form (ScaledParams::skip_rc_exact), else IEEE division
template <bool SRC>
__device__ __forceinline__ float div_skip(float num, float skip, float skip_rc) {
  return SRC ? div_rc(num, skip, skip_rc) : __fdiv_rn(num, skip);
}

template <int NX, bool UNIFORM, bool SRC>
__device__ __forceinline__ void window_rows_bayer(const ScaledParams &p, const float ax[NX], int from_x, int nx, int from_y, int to_y,
                                                  float center_y, bool g00, F2 &g, F2 &x0, F2 &x1) {
  const float range = p.range, rc = p.range_rc, skip_y = p.skip_y, sub_a = p.sub_a, sub_b = p.sub_b;
  const uint16_t *rowp = p.raw + (long long)(from_y + p.crop_y - p.src_row0) * p.raw_pitch + p.crop_x + from_x;
  for (int y = from_y; y <= to_y; y += 2, rowp += 2 * p.raw_pitch) {
    // both rows of a pair are loaded before either is used (ten loads in flight); the last pair of a window with an odd
    // number of rows loads its only row twice and skips the second half of the arithmetic
    const bool two = y + 1 <= to_y;
    const uint16_t *rowq = two ? rowp + p.raw_pitch : rowp;
    uint32_t s0[NX], s1[NX];
#pragma unroll
    for (int k = 0; k < NX; k++) s0[k] = __ldg(rowp + (UNIFORM ? k : min(k, nx - 1)));
#pragma unroll
    for (int k = 0; k < NX; k++) s1[k] = __ldg(rowq + (UNIFORM ? k : min(k, nx - 1)));
    const float d0 = div_skip<SRC>((float)y - center_y, skip_y, p.skip_y_rc);
    bayer_row<NX, UNIFORM, 0>(s0, ax, d0 * d0, nx, g00, sub_a, sub_b, range, rc, g, x0);
    if (two) {
      const float d1 = div_skip<SRC>((float)(y + 1) - center_y, skip_y, p.skip_y_rc);
      bayer_row<NX, UNIFORM, 1>(s1, ax, d1 * d1, nx, g00, sub_a, sub_b, range, rc, g, x1);
    }
  }
}

template <int NX, bool UNIFORM, bool SRC>
__device__ __forceinline__ void window_taps_bayer(const ScaledParams &p, const CfaDev &cfa, int from_x, int nx, int from_y,
                                                  int to_y, float center_x, float center_y, F2 acc[4]) {
  float ax[NX];
#pragma unroll
  for (int k = 0; k < NX; k++) {
    const float delta_x = div_skip<SRC>((float)(from_x + k) - center_x, p.skip_x, p.skip_x_rc);
    ax[k] = 1.0f - (delta_x * delta_x);
  }
  const int py = from_y & 1, pxb = from_x & 1;
  const bool g00 = cfa.pat[py * 48 + pxb] == 1;
  F2 g{0.f, 0.f}, x0{0.f, 0.f}, x1{0.f, 0.f};
  window_rows_bayer<NX, UNIFORM, SRC>(p, ax, from_x, nx, from_y, to_y, center_y, g00, g, x0, x1);
  // the non-green colour of the window's first row (0 = red or 2 = blue); the second row holds the other one
  const int c0 = g00 ? cfa.pat[py * 48 + (pxb ^ 1)] : cfa.pat[py * 48 + pxb];
  acc[1] = g;
  acc[0] = c0 == 0 ? x0 : x1;
  acc[2] = c0 == 0 ? x1 : x0;
}

// Window of output pixel (row, col): scaling.rs:77-89 with topleft = (0,0), skip_x_y = skip_y_x = 0, every term of the
// reference's expressions kept (0.0 * x included: it decides the sign of a zero and propagates NaN like the reference)
struct ScaledWindow {
  int from_x, to_x, from_y, to_y;
  float center_x, center_y;
};
__device__ __forceinline__ ScaledWindow scaled_window(const ScaledParams &p, int row, int col) {
  const float frow = (float)row, frow1 = (float)(row + 1), fcol = (float)col, fcol1 = (float)(col + 1);
  const float rfrom_x = 0.0f + 0.0f * frow;
  const float rto_x = 0.0f + 0.0f * frow1;
  const float rfrom_y = 0.0f + p.skip_y * frow;
  const float rto_y = 0.0f + p.skip_y * frow1;
  const float rcenter_x = 0.0f + (0.0f * frow) + __fdiv_rn(0.0f, 2.0f) - 0.5f;
  const float rcenter_y = 0.0f + (p.skip_y * frow) + __fdiv_rn(p.skip_y, 2.0f) - 0.5f;
  ScaledWindow w;
  w.from_x = min(p.width - 1, f2i_sat(floorf(rfrom_x + (p.skip_x * fcol))));
  w.to_x = min(p.width - 1, f2i_sat(floorf(rto_x + (p.skip_x * fcol1))));
  w.from_y = min(p.height - 1, f2i_sat(floorf(rfrom_y + (0.0f * fcol))));
  w.to_y = min(p.height - 1, f2i_sat(floorf(rto_y + (0.0f * fcol1))));
  w.center_x = rcenter_x + (p.skip_x * fcol) + __fdiv_rn(p.skip_x, 2.0f);
  w.center_y = rcenter_y + (0.0f * fcol) + __fdiv_rn(0.0f, 2.0f);
  return w;
}

// scaled_demosaic of ONE output pixel of an RGB Bayer frame (scaling.rs:76-127 with the CFA binning of :109-112): the
// window's taps, then sums[c] / counts[c] (IEEE division; a colour without weight stays 0.0, :120-126).  Whole warps
// call this together (warp-wide votes pick the unrolled loop).  Needs p.exact_rc and windows of at most kMaxCols columns.
__device__ __forceinline__ void scaled_pixel_bayer(const ScaledParams &p, const CfaDev &cfa, int row, int col, float px[3]) {
  const ScaledWindow w = scaled_window(p, row, col);
  const int nx = w.to_x - w.from_x + 1;
  F2 acc[4] = {F2{0.f, 0.f}, F2{0.f, 0.f}, F2{0.f, 0.f}, F2{0.f, 0.f}};  // {sums[c], counts[c]}
  const int nx_max = __reduce_max_sync(0xffffffffu, nx), nx_min = __reduce_min_sync(0xffffffffu, nx);
  if (nx_max == 5 && nx_min == 5) {
    if (p.skip_rc_exact) window_taps_bayer<5, true, true>(p, cfa, w.from_x, nx, w.from_y, w.to_y, w.center_x, w.center_y, acc);
    else window_taps_bayer<5, true, false>(p, cfa, w.from_x, nx, w.from_y, w.to_y, w.center_x, w.center_y, acc);
  } else if (nx_max <= 6) {
    window_taps_bayer<6, false, false>(p, cfa, w.from_x, nx, w.from_y, w.to_y, w.center_x, w.center_y, acc);
  } else {
    window_taps_bayer<kMaxCols, false, false>(p, cfa, w.from_x, nx, w.from_y, w.to_y, w.center_x, w.center_y, acc);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) px[k] = acc[k].y > 0.0f ? __fdiv_rn(acc[k].x, acc[k].y) : 0.0f;
}

// Two output pixels of one column in consecutive rows (row, row + 1): their windows share the column geometry — from_x,
// to_x, center_x and with them 1 - delta_x^2 of every window column do not depend on the row (scaling.rs:77-98 with
// skip_x_y == 0) — so that part is computed once.  `two` false: only (row, col) exists; pb is a copy of pa.
template <int NX, bool UNIFORM, bool SRC>
__device__ __forceinline__ void pair_taps_bayer(const ScaledParams &p, const CfaDev &cfa, const ScaledWindow &wa, const ScaledWindow &wb,
                                                int nx, bool two, F2 acc_a[3], F2 acc_b[3]) {
  float ax[NX];
#pragma unroll
  for (int k = 0; k < NX; k++) {
    const float delta_x = div_skip<SRC>((float)(wa.from_x + k) - wa.center_x, p.skip_x, p.skip_x_rc);
    ax[k] = 1.0f - (delta_x * delta_x);
  }
  const int pxb = wa.from_x & 1;
  auto one = [&](const ScaledWindow &w, F2 acc[3]) {
    const int py = w.from_y & 1;
    const bool g00 = cfa.pat[py * 48 + pxb] == 1;
    F2 g{0.f, 0.f}, x0{0.f, 0.f}, x1{0.f, 0.f};
    window_rows_bayer<NX, UNIFORM, SRC>(p, ax, w.from_x, nx, w.from_y, w.to_y, w.center_y, g00, g, x0, x1);
    const int c0 = g00 ? cfa.pat[py * 48 + (pxb ^ 1)] : cfa.pat[py * 48 + pxb];
    acc[1] = g;
    acc[0] = c0 == 0 ? x0 : x1;
    acc[2] = c0 == 0 ? x1 : x0;
  };
  one(wa, acc_a);
  if (two) one(wb, acc_b);
  else { acc_b[0] = acc_a[0]; acc_b[1] = acc_a[1]; acc_b[2] = acc_a[2]; }
}
__device__ __forceinline__ void scaled_pair_bayer(const ScaledParams &p, const CfaDev &cfa, int row, int col, bool two, float pa[3],
                                                  float pb[3]) {
  const ScaledWindow wa = scaled_window(p, row, col);
  ScaledWindow wb = wa;
  if (two) {   // the row part of the window of (row + 1, col); the column part is wa's
    const ScaledWindow t = scaled_window(p, row + 1, col);
    wb.from_y = t.from_y; wb.to_y = t.to_y; wb.center_y = t.center_y;
  }
  const int nx = wa.to_x - wa.from_x + 1;
  F2 acc_a[3], acc_b[3];
  const int nx_max = __reduce_max_sync(0xffffffffu, nx), nx_min = __reduce_min_sync(0xffffffffu, nx);
  if (nx_max == 5 && nx_min == 5) {
    if (p.skip_rc_exact) pair_taps_bayer<5, true, true>(p, cfa, wa, wb, nx, two, acc_a, acc_b);
    else pair_taps_bayer<5, true, false>(p, cfa, wa, wb, nx, two, acc_a, acc_b);
  } else if (nx_max <= 6) {
    pair_taps_bayer<6, false, false>(p, cfa, wa, wb, nx, two, acc_a, acc_b);
  } else {
    pair_taps_bayer<kMaxCols, false, false>(p, cfa, wa, wb, nx, two, acc_a, acc_b);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    pa[k] = acc_a[k].y > 0.0f ? __fdiv_rn(acc_a[k].x, acc_a[k].y) : 0.0f;
    pb[k] = acc_b[k].y > 0.0f ? __fdiv_rn(acc_b[k].x, acc_b[k].y) : 0.0f;
  }
}

// host (ipb_fused.cu): the geometry / level-mapping part of ScaledParams from the launch arguments, including the
// exhaustive check behind skip_rc_exact
void fill_scaled_params(const FusedArgs &a, const CfaDev &cfa, ScaledParams *p);
// host: the check behind ScaledParams::skip_rc_exact for a frame geometry (cropped frame -> output size)
bool scaled_skip_division_exact(size_t width, size_t height, size_t nwidth, size_t nheight);

}  // namespace ipb
