// ipb_spec_host.cu — host side of the speculative 8-bit kernel: its two tables, its folded constants, and the
// certified bound `delta` on |cheap linear value - reference linear value| that decides which pixels are recomputed.
// Pure host arithmetic (double precision), no device code.  DESIGN.md "speculative pass" carries the derivation in prose;
// every term below names the roundings it covers.  u = 2^-24 is the unit round-off of f32.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ipb_spec.h"

namespace ipb {

namespace {

constexpr double U = 5.9604644775390625e-08;            // 2^-24
const double kE = (double)(216.0f / 24389.0f);           // color_conversions.rs:121 (f32 value)
const double kK = (double)(24389.0f / 27.0f);            // :122
const double kWhite[3] = {(double)0.95047f, 1.0, (double)1.08883f};
constexpr double kLutStep = 1.0 / 8191.0;

// SplineFunc::interpolate (curves.rs:126-157) as a real function of a real argument (f32 coefficients)
double spline_real(const SplineDev &s, double v) {
  if (s.n == 0) return v;
  const int last = s.n - 1;
  if (v >= (double)s.x[last]) return (double)s.y[last];
  if (v <= (double)s.x[0]) return (double)s.y[0];
  int i = 0;
  while (i + 1 < s.nseg && v >= (double)s.x[i + 1]) i++;
  const double d = v - (double)s.x[i];
  return (double)s.y[i] + (double)s.c1[i] * d + (double)s.c2[i] * d * d + (double)s.c3[i] * d * d * d;
}
// sum of the magnitudes of the cubic's terms at v: scales the reference's own rounding error in the evaluation
double spline_mag(const SplineDev &s, double v) {
  if (s.n == 0) return 0.0;
  const int last = s.n - 1;
  if (v >= (double)s.x[last] || v <= (double)s.x[0]) return 0.0;
  int i = 0;
  while (i + 1 < s.nseg && v >= (double)s.x[i + 1]) i++;
  const double d = fabs(v - (double)s.x[i]);
  return fabs((double)s.y[i]) + 2.0 * fabs((double)s.c1[i]) * d + 4.0 * fabs((double)s.c2[i]) * d * d +
         6.0 * fabs((double)s.c3[i]) * d * d * d;
}
// the basecurve in f-space: fy -> fy'
double S_real(const SplineDev &s, double f) { return (100.0 * spline_real(s, (116.0 * f - 16.0) / 100.0) + 16.0) / 116.0; }

double lab_f(double v) { return v > kE ? cbrt(v) : (kK * v + 16.0) / 116.0; }
double lab_fp(double v) { return v > kE ? 1.0 / (3.0 * cbrt(v) * cbrt(v)) : kK / 116.0; }
double lab_g(double t) { return t > 6.0 / 29.0 ? t * t * t : (116.0 * t - 16.0) / kK; }
double lab_gp(double t) { return t > 6.0 / 29.0 ? 3.0 * t * t : 116.0 / kK; }
// |table lerp - f| of the reference's 8193-entry XYZ_LAB_TRANSFORM (color_conversions.rs:80-115) inside [0, 1]: chord
// error h^2/8 * |f''| with f'' taken at the low end of the segment (|f''| decreases), zero on the linear part except in
// the segment that contains the joint
double lab_gap(double v) {
  if (v < kE - kLutStep || v > 1.0) return 0.0;
  const double a = std::max(v - kLutStep, kE);
  return kLutStep * kLutStep / 8.0 * (2.0 / 9.0) * pow(a, -5.0 / 3.0);
}

}  // namespace

bool spec_build(const ColorParams &P, float black, float range, float mufu_rel_err, float delta_override,
                const std::vector<float> &thresholds, std::vector<uint32_t> *g8a, std::vector<float2> *stab,
                SpecParams *consts, float delta_out[4]) {
  if (thresholds.size() != 255 || P.use_e || P.linear) return false;
  SpecParams c;
  memset(&c, 0, sizeof(c));
  const SplineDev &sp = P.sp;

  // ------------------------------------------------------------ folded constants
  for (int j = 0; j < 3; j++)
    if (!(P.mul[j] > 0.0f)) return false;  // a zero or negative gain cannot be folded into a clip limit
  if (P.mul[1] != 1.0f) return false;      // normalize_wbs divides by the green gain
  c.lim_r = 1.0f / P.mul[0];
  c.lim_b = 1.0f / P.mul[2];
  double M[3][3], RO[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      M[i][j] = (double)P.cm[i * 4 + j] * (double)P.mul[j] / kWhite[i];
      RO[i][j] = (double)P.rgbm[i * 3 + j] * kWhite[j];
      c.m[i][j] = (float)M[i][j];
      c.ro[i][j] = (float)RO[i][j];
    }
  const double f0 = 16.0 / 116.0, h = (1.0 - f0) / kSpecSTabN;
  c.s_scale = (float)(1.0 / (h * (kSpecSTabN + 2)));
  c.s_off = (float)((1.0 - f0 / h) / (kSpecSTabN + 2));
  c.bias58 = 0x58000000u;

  // ------------------------------------------------------------ reachable ranges
  // demosaiced samples lie in [gmin, 1] (gofloat.rs:127 clips at 1 only), clipped white-balanced ones in [m*gmin, min(m,1)]
  const double gmin = std::min(0.0, -(double)black / (double)range), gmax = 1.0;
  double xlo[3], xhi[3], neg[3];
  for (int i = 0; i < 3; i++) {
    xlo[i] = xhi[i] = neg[i] = 0.0;
    for (int j = 0; j < 3; j++) {
      const double m = (double)P.mul[j], coef = (double)P.cm[i * 4 + j] / kWhite[i];
      const double cmax = std::min(m * gmax, 1.0), cmin = m * gmin;  // cmin <= 0 <= cmax
      xhi[i] += coef > 0 ? coef * cmax : coef * cmin;
      xlo[i] += coef > 0 ? coef * cmin : coef * cmax;
      neg[i] += coef > 0 ? coef * -cmin : -coef * cmax;  // largest possible sum of the negative terms
    }
  }

  for (int i = 0; i < 3; i++)
    if (xhi[i] > 3.9 || xlo[i] < -2.0) return false;  // outside what the XU-pipe cube root was measured on

  // ------------------------------------------------------------ basecurve table S(fy) as {intercept, slope} chords
  stab->assign(kSpecSTabEntries, make_float2(0.0f, 0.0f));
  const double s_lo = S_real(sp, f0), s_hi = S_real(sp, 1.0);
  std::vector<double> A(kSpecSTabEntries), B(kSpecSTabEntries);
  for (int k = 0; k < kSpecSTabEntries; k++) {
    if (sp.n == 0) { A[k] = 0.0; B[k] = 1.0; }
    else if (k == 0) { A[k] = s_lo; B[k] = 0.0; }
    else if (k > kSpecSTabN) { A[k] = s_hi; B[k] = 0.0; }
    else {
      const double fa = f0 + (k - 1) * h, fb = f0 + k * h;
      const double sa = S_real(sp, fa), sb = S_real(sp, fb);
      B[k] = (sb - sa) / h;
      A[k] = sa - B[k] * fa;
    }
    (*stab)[k] = make_float2((float)A[k], (float)B[k]);
  }
  // largest deviation of the table, as the kernel evaluates it, from S over the reachable fy range: chord error,
  // coefficient rounding, the kinks at the curve's end points and knots, and a segment chosen one off at a boundary
  const double fy_lo = lab_f(std::max(xlo[1], (double)kSpecYMin)) - 1e-3, fy_hi = lab_f(xhi[1]) + 1e-3;
  auto tab_eval = [&](double f, int shift) {
    double pos = (f - f0) / h + 1.0;
    int k = (int)floor(std::min(std::max(pos, 0.0), (double)(kSpecSTabN + 2)));
    k = std::min(std::max(k + shift, 0), kSpecSTabN + 2);
    return (double)(*stab)[k].x + (double)(*stab)[k].y * f;
  };
  double chord = 0.0, smag = 1.0, ls_max = 0.0, sp_max = 0.0, d_max = 0.0;
  {
    std::vector<double> pts;
    const int nsamp = 16 * (kSpecSTabN + 2);
    for (int i = 0; i <= nsamp; i++) pts.push_back(fy_lo + (fy_hi - fy_lo) * i / nsamp);
    for (int i = 0; i < sp.n; i++) {
      const double fk = (100.0 * (double)sp.x[i] + 16.0) / 116.0;
      for (double eps : {-1e-9, 0.0, 1e-9}) pts.push_back(fk + eps);
    }
    for (double f : pts) {
      const double s = S_real(sp, f);
      chord = std::max(chord, fabs(tab_eval(f, 0) - s));
      smag = std::max(smag, fabs(s));
      d_max = std::max(d_max, s - f);
      const double pos = (f - f0) / h + 1.0, frac = pos - floor(pos);
      (void)frac;
    }
    // the kernel's index comes from one f32 FMA: it can be one off within 3e-4 of a segment boundary (2 ulp of the
    // scaled position), where it then extrapolates the neighbouring chord
    for (int k = 1; k <= kSpecSTabN + 1; k++)
      for (double off : {-3e-4, 3e-4}) {
        const double f = f0 + (k - 1 + off) * h;
        if (f < fy_lo || f > fy_hi) continue;
        chord = std::max(chord, fabs(tab_eval(f, off < 0 ? +1 : -1) - S_real(sp, f)));
      }
    for (int k = 0; k <= kSpecSTabN + 2; k++) {
      ls_max = std::max(ls_max, fabs(B[k] - 1.0));
      sp_max = std::max(sp_max, fabs(B[k]));
    }
    if (sp.n == 0) { ls_max = 0.0; sp_max = 1.0; }
  }
  double amax = 0.0;
  for (int k = 0; k <= kSpecSTabN + 2; k++) amax = std::max(amax, fabs(A[k]));
  // the kernel's fma(fy, slope, intercept): rounded coefficients and one rounding of the result
  const double e_stab = chord + U * (fabs(smag) + sp_max * std::max(fabs(fy_lo), fabs(fy_hi)) + amax);

  // ------------------------------------------------------------ the bound
  // Cheap (A) and reference (E) start from IDENTICAL demosaiced samples: level mapping uses the verified exact
  // reciprocal form and the Bayer means add in the reference's tap order.  From there the two computations are followed
  // over a grid of the clipped, white-balanced colour cube [m*gmin, min(m,1)]^3 (denser towards the dark end, where
  // the transfer function bends); at every grid point the first-order error terms of both are added up:
  // (1) XYZ ratios.  E: c*mul (u), three products (u each), two sums (u each), the division (u): <= 5u * sum|terms|.
  //     A: folded coefficients (1.5u), rounded clip limit (u), three FMA roundings: <= 5u * sum|terms|.
  // (2) Lab transfer function.  E: table lerp (gap to the true function + 3.5u*f), libm cbrtf above 1 (2u*f), the line
  //     below 0 (3u); A: XU pipe (measured relative error over every float), the line (3u).  Input error times f'.
  // (3) basecurve and the a / b legs in f-space.
  //     E, L leg: 116*fy (u), -16 (u), /100 (u), spline (term magnitudes * u, three sums), *100 (u), +16 (u), /116 (u).
  //     E, a / b leg: fx-fy (u), *500 (u), +127 (u), /255 (u), *255 (u), -127 (u), /500 (u), +fy' (u) on magnitudes
  //     <= 127 + 500|d|, d = fx - fy.   A: table (e_stab), fy' - fy (u), fx + D (u).
  // (4) inverse transfer function: error in t times g'(t), two products (u each) on both sides.
  // (5) output matrix: E 5u, A 4u per term; input errors times |coefficient|.
  const double mufu = std::max((double)mufu_rel_err * 1.25, 4.0 * U);
  std::vector<double> axis[3];
  for (int j = 0; j < 3; j++) {
    const double m = (double)P.mul[j], cmax = std::min(m * gmax, 1.0), cmin = m * gmin;
    for (int k = 0; k <= 4; k++) axis[j].push_back(cmin * (4 - k) / 4.0);                  // cmin .. 0
    for (int k = 0; k < 12; k++) axis[j].push_back(cmax * pow(10.0, -4.0 + 3.0 * k / 12.0));  // 1e-4*cmax .. 0.1*cmax
    for (int k = 0; k <= 27; k++) axis[j].push_back(cmax * (0.1 + 0.9 * k / 27.0));          // 0.1*cmax .. cmax
  }
  const double hS = h * 0.5;
  double delta = 0.0, dch[3] = {0, 0, 0}, worst_ex[3] = {0, 0, 0};
  for (double ca : axis[0])
    for (double cb : axis[1])
      for (double cc : axis[2]) {
        const double cv[3] = {ca, cb, cc};
        double x[3], f[3], ef[3];
        for (int i = 0; i < 3; i++) {
          double sum = 0.0, mag = 0.0;
          for (int j = 0; j < 3; j++) {
            const double term = (double)P.cm[i * 4 + j] / kWhite[i] * cv[j];
            sum += term;
            mag += fabs(term);
          }
          x[i] = sum;
          const double ex = 10.0 * U * mag;
          f[i] = lab_f(sum);
          ef[i] = lab_fp(sum - ex - 0.05 * fabs(sum)) * ex + lab_gap(sum) + (3.5 * U + mufu) * fabs(f[i]) +
                  (sum < kE + 2.0 * kLutStep ? 6.0 * U : 0.0);
        }
        if (x[1] < kSpecYMin - 1e-4) continue;  // the kernel recomputes pixels below the certified domain
        const double fy = f[1], Sy = S_real(sp, fy);
        const double sl1 = fabs(S_real(sp, fy + hS) - Sy) / hS, sl2 = fabs(Sy - S_real(sp, fy - hS)) / hS;
        const double spl = sp.n ? std::max(sl1, sl2) : 1.0;
        const double ls = sp.n ? std::max(fabs(sl1 - 1.0), fabs(sl2 - 1.0)) : 0.0;
        const double l = (116.0 * fy - 16.0) / 100.0, sv = spline_real(sp, l);
        const double e_l = U * (116.0 * fabs(fy) + fabs(116.0 * fy - 16.0)) / 100.0 + U * fabs(l);
        const double e_sp = sp.n ? U * (spline_mag(sp, l) + 3.0 * fabs(sv)) : 0.0;
        const double e_sE = (100.0 * (spl * e_l + e_sp) + U * (100.0 * fabs(sv) + 116.0 * fabs(Sy))) / 116.0 + U * fabs(Sy);
        const double e_fyp = e_sE + e_stab;
        const double D = Sy - fy;
        // Error of the three cube-root-domain values t = (fx + D, fy', fz + D), cheap minus reference:
        //   dt_x = dfx + ind_x + cS + (S' - 1) dfy,   dt_y = cS + S' dfy,   dt_z = dfz + ind_z + cS + (S' - 1) dfy
        // dfx, dfy, dfz: the transfer-function errors (|.| <= ef); ind: the roundings of the a / b legs; cS: the error of the
        // basecurve step itself (|cS| <= e_fyp) — ONE number for all three, because both computations add the same fy'
        // to the a / b differences; S': a secant slope of S near fy (between the two one-sided slopes).  The common terms
        // reach an output channel through the SIGNED sum of its matrix row times g'(t), which is much smaller than the
        // sum of magnitudes (the rows of the XYZ -> RGB matrix sum to about 1 but hold entries up to 3.2).
        double EXp[3], Xp[3], gpm[3], gpc[3], dev[3], eind[3], tj[3];
        for (int i = 0; i < 3; i++) {
          double et;
          if (i == 1) {
            et = spl * ef[1] + e_fyp;
            tj[i] = Sy;
            eind[i] = 0.0;
          } else {
            const double e_ab = U * (2.0 + 7.1 * fabs(f[i] - fy)) + 1.5 * U;
            eind[i] = ef[i] + e_ab + 2.0 * U * (fabs(f[i]) + fabs(D));
            et = eind[i] + ls * ef[1] + e_fyp;
            tj[i] = f[i] + D;
          }
          gpc[i] = lab_gp(tj[i]);
          gpm[i] = std::max(lab_gp(tj[i] + et), lab_gp(tj[i] - et));
          dev[i] = std::max(fabs(lab_gp(tj[i] + et) - gpc[i]), fabs(lab_gp(tj[i] - et) - gpc[i]));
          EXp[i] = gpm[i] * et + 4.0 * U * fabs(lab_g(tj[i])) + 4.0 * U * 16.0 / kK;   // magnitude bound (sum of everything)
          Xp[i] = fabs(lab_g(tj[i])) + EXp[i];
          worst_ex[i] = std::max(worst_ex[i], EXp[i]);
        }
        const double slopes[2] = {sp.n ? sl1 : 1.0, sp.n ? sl2 : 1.0};
        for (int ch = 0; ch < 3; ch++) {
          double d = 0.0, sum_c = 0.0, sum_dev = 0.0;
          for (int j = 0; j < 3; j++) {
            d += fabs(RO[ch][j]) * (gpm[j] * eind[j] + 4.0 * U * fabs(lab_g(tj[j])) + 4.0 * U * 16.0 / kK + 9.0 * U * Xp[j]);
            sum_c += RO[ch][j] * gpc[j];
            sum_dev += fabs(RO[ch][j]) * dev[j];
          }
          d += e_fyp * (fabs(sum_c) + sum_dev);
          double worst_y = 0.0;
          for (double sl : slopes) {
            const double k[3] = {sl - 1.0, sl, sl - 1.0};
            double sk = 0.0, skd = 0.0;
            for (int j = 0; j < 3; j++) {
              sk += RO[ch][j] * gpc[j] * k[j];
              skd += fabs(RO[ch][j]) * dev[j] * fabs(k[j]);
            }
            worst_y = std::max(worst_y, fabs(sk) + skd);
          }
          d += ef[1] * worst_y;
          if (d > delta && getenv("IPB_SPEC_DEBUG2"))
            fprintf(stderr, "  ch %d c (%.4f %.4f %.4f) x (%.4f %.4f %.4f) ef %.1f %.1f %.1f u e_sE %.1f e_stab %.1f ls %.3f spl %.3f D %.3f EX %.1f %.1f %.1f u X' %.3f %.3f %.3f d %.3g\n",
                    ch, ca, cb, cc, x[0], x[1], x[2], ef[0] / U, ef[1] / U, ef[2] / U, e_sE / U, e_stab / U, ls, spl, D, EXp[0] / U, EXp[1] / U, EXp[2] / U, Xp[0], Xp[1], Xp[2], d);
          delta = std::max(delta, d);
          dch[ch] = std::max(dch[ch], d);
        }
      }
  delta *= 1.25;  // variation between grid points, second-order terms
  for (int ch = 0; ch < 3; ch++) dch[ch] *= 1.25;
  if (getenv("IPB_SPEC_DEBUG"))
    fprintf(stderr, "spec_build: x in [%.3f,%.3f] [%.3f,%.3f] [%.3f,%.3f] chord %.3g e_stab %.3g u  worst EX %.3g %.3g %.3g  mufu %.3g  delta %.4g\n",
            xlo[0], xhi[0], xlo[1], xhi[1], xlo[2], xhi[2], chord, e_stab / U, worst_ex[0], worst_ex[1], worst_ex[2], mufu, delta);

  if (delta_override > 0.0f) delta = dch[0] = dch[1] = dch[2] = (double)delta_override;
  if (!(delta > 0.0) || delta > 8.0e-5) return false;  // thresholds are >= 3.0e-4 (2536 F units) apart: at most one per
                                                        // extended segment of 1024 + 2 * (2 * deltaF + 4) units

  // ------------------------------------------------------------ gamma table in fixed point
  // F = round(v * kSpecFScale) is what the kernel reads from the bit pattern of 1 + v * (1 - 2^-13); a reference
  // threshold T sits at tau = T * kSpecFScale.  With |v_cheap - v_ref| <= delta and half a unit of rounding in F,
  // |F - tau| > deltaF = ceil(delta * 2^23) + 2 on every threshold means both values lie on the same side of all of them.
  uint32_t dFc[3], dF = 0;
  for (int ch = 0; ch < 3; ch++) {
    dFc[ch] = (uint32_t)ceil(dch[ch] * 8388608.0) + 2u;
    dF = std::max(dF, dFc[ch]);
  }
  const int64_t margin = (int64_t)dF + 4;  // thresholds within deltaF of an F of the segment; the index may be one over
  g8a->assign(kSpecG8Entries, 0u);
  std::vector<double> tau(255);
  for (int i = 0; i < 255; i++) tau[i] = (double)thresholds[i] * kSpecFScale;
  for (int k = 0; k < kSpecG8Entries; k++) {
    const int64_t lo = (int64_t)k * 1024 - margin, hi = (int64_t)k * 1024 + 1023 + margin;
    int found = -1, below = 0;
    for (int i = 0; i < 255; i++) {
      if (tau[i] < (double)lo) below++;
      else if (tau[i] <= (double)hi) { if (found >= 0) return false; found = i; }
    }
    // entry + bits(u_c) = (base << 24) + 2^24 + (F + deltaF_c - thrF): the byte is base below the threshold and base + 1
    // from it on.  A segment without a threshold pretends to one 2^19 units below its start (byte = base + 1 throughout):
    // its distance field stays near 2^19, which the weighted comparison (products below 2^32) never takes for "close".
    uint32_t base, thrF;
    if (found >= 0) { base = (uint32_t)found; thrF = (uint32_t)ceil(tau[found]); }
    else { base = ((uint32_t)below - 1u) & 0xffu; thrF = (uint32_t)k * 1024u - 524288u; }
    (*g8a)[k] = (base << 24) + (16777216u - thrF) - 0x3F800000u;
  }
  for (int ch = 0; ch < 3; ch++) {
    c.one[ch] = 1.0f + (float)dFc[ch] * 1.1920928955078125e-07f;  // exact: deltaF < 2^10
    c.wmul[ch] = 256u * ((4u * dF) / dFc[ch]);
  }
  c.amb_t = 256u * 4u * 2u * dF;
  c.y_min = kSpecYMin;
  *consts = c;
  delta_out[0] = (float)delta;
  for (int ch = 0; ch < 3; ch++) delta_out[1 + ch] = (float)dch[ch];
  return true;
}

}  // namespace ipb
