/*
 * lanczos.c — CPU statement of the Lanczos-a separable resampler (EXTENSION, test infrastructure).
 *
 * PARITY UNPINNED: the reference has no Lanczos resampler — only the FIXME at src/scaling.rs:101-103 ("A good windowed
 * sinc function like Lanczos ...").  The task's north star names "Lanczos scaling", so the product ships one as an
 * optional op (ipb_lanczos_resize) and this file is its independent CPU statement: same definition, same order of
 * f32 operations, checked bit for bit by tests/test_gpu_lanczos.py and by known answers in tests/test_lanczos_cpu.py
 * (constant images, impulse response, weights summing to one, symmetry).
 *
 * Definition (the usual separable area-aware convolution): for one axis with n_in samples resampled to n_out,
 *   scale = n_in / n_out, fscale = max(scale, 1), support = a * fscale,
 *   centre(i) = (i + 0.5) * scale, taps x in [max(0, trunc(centre - support + 0.5)), min(n_in, trunc(centre + support + 0.5))),
 *   w(x) = L((x - centre + 0.5) / fscale), L(t) = sinc(t) * sinc(t / a) for |t| < a else 0, sinc(t) = sin(pi t)/(pi t),
 * weights computed in double, normalised to sum 1, rounded to f32.  Rows first (horizontal pass into an f32
 * intermediate), then columns.  A sample is acc = 0; acc += w[k] * in[k] for ascending k, in f32, no FMA.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static double lanczos_kernel(double t, int a) {
  if (t < 0) t = -t;
  if (t >= (double)a) return 0.0;
  if (t == 0.0) return 1.0;
  const double pt = M_PI * t;
  return (sin(pt) / pt) * (sin(pt / (double)a) / (pt / (double)a));
}

/* weights of one axis: start[i], count[i], w[i * ksize + k]; returns ksize */
static size_t lanczos_weights(size_t n_in, size_t n_out, int a, int **start, int **count, float **w) {
  const double scale = (double)n_in / (double)n_out;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = (double)a * fscale;
  const size_t ksize = (size_t)ceil(support) * 2 + 1;
  *start = (int *)malloc(n_out * sizeof(int));
  *count = (int *)malloc(n_out * sizeof(int));
  *w = (float *)calloc(n_out * ksize, sizeof(float));
  double *tmp = (double *)malloc(ksize * sizeof(double));
  for (size_t i = 0; i < n_out; i++) {
    const double centre = ((double)i + 0.5) * scale;
    long xmin = (long)(centre - support + 0.5);
    long xmax = (long)(centre + support + 0.5);
    if (xmin < 0) xmin = 0;
    if (xmax > (long)n_in) xmax = (long)n_in;
    const long n = xmax - xmin;
    double sum = 0.0;
    for (long k = 0; k < n; k++) {
      tmp[k] = lanczos_kernel(((double)(xmin + k) - centre + 0.5) / fscale, a);
      sum += tmp[k];
    }
    for (long k = 0; k < n; k++) (*w)[i * ksize + (size_t)k] = (float)(sum != 0.0 ? tmp[k] / sum : tmp[k]);
    (*start)[i] = (int)xmin;
    (*count)[i] = (int)n;
  }
  free(tmp);
  return ksize;
}

void orc_lanczos_weights(size_t n_in, size_t n_out, int a, int *start, int *count, float *w, size_t *ksize_out) {
  int *s, *c;
  float *ww;
  const size_t ks = lanczos_weights(n_in, n_out, a, &s, &c, &ww);
  if (ksize_out) *ksize_out = ks;
  if (start) memcpy(start, s, n_out * sizeof(int));
  if (count) memcpy(count, c, n_out * sizeof(int));
  if (w) memcpy(w, ww, n_out * ks * sizeof(float));
  free(s); free(c); free(ww);
}

orc_buffer *orc_lanczos_resize(const orc_buffer *buf, size_t nw, size_t nh, int a) {
  if (!buf || nw == 0 || nh == 0 || a < 1 || a > 8 || buf->width == 0 || buf->height == 0) return NULL;
  const size_t W = buf->width, H = buf->height, C = buf->colors;
  int *sx, *cx, *sy, *cy;
  float *wx, *wy;
  const size_t kx = lanczos_weights(W, nw, a, &sx, &cx, &wx);
  const size_t ky = lanczos_weights(H, nh, a, &sy, &cy, &wy);
  float *mid = (float *)malloc(H * nw * C * sizeof(float));
  orc_buffer *out = orc_buffer_new(nw, nh, C, buf->monochrome);
#pragma omp parallel for
  for (long y = 0; y < (long)H; y++)
    for (size_t x = 0; x < nw; x++)
      for (size_t c = 0; c < C; c++) {
        float acc = 0.0f;
        for (int k = 0; k < cx[x]; k++) acc = acc + wx[x * kx + (size_t)k] * buf->data[((size_t)y * W + (size_t)(sx[x] + k)) * C + c];
        mid[((size_t)y * nw + x) * C + c] = acc;
      }
#pragma omp parallel for
  for (long y = 0; y < (long)nh; y++)
    for (size_t x = 0; x < nw; x++)
      for (size_t c = 0; c < C; c++) {
        float acc = 0.0f;
        for (int k = 0; k < cy[y]; k++) acc = acc + wy[(size_t)y * ky + (size_t)k] * mid[((size_t)(sy[y] + k) * nw + x) * C + c];
        out->data[((size_t)y * nw + x) * C + c] = acc;
      }
  free(mid); free(sx); free(cx); free(wx); free(sy); free(cy); free(wy);
  return out;
}
