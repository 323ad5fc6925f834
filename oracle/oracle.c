/*
 * oracle.c — CPU restatement of the pedrocr/imagepipe OpBuffer hot path (see oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Build: see oracle/Makefile (-O2 -ffp-contract=off -fopenmp).
 * Citations are to /root/reference/<file>:<lines>.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ helpers */

/* Rust `f as usize` / `f as isize` saturate and map NaN to 0. */
static inline size_t f2usize(float f) {
  if (!(f > 0.0f)) return 0; /* negatives and NaN */
  if (f >= 18446744073709551616.0f) return (size_t)-1;
  return (size_t)f;
}
static inline long f2isize(float f) {
  if (f != f) return 0;
  if (f >= 9223372036854775808.0f) return 0x7fffffffffffffffL;
  if (f <= -9223372036854775808.0f) return (long)0x8000000000000000UL;
  return (long)f;
}
static inline size_t umin(size_t a, size_t b) { return a < b ? a : b; }

static int g_threads = 0;
void orc_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  else omp_set_num_threads(omp_get_num_procs());
#endif
}
int orc_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ------------------------------------------------------------------ buffer.rs */

/* src/buffer.rs:24-32 OpBuffer::new — zero-filled (vec![0.0; n] is a calloc) */
orc_buffer *orc_buffer_new(size_t w, size_t h, size_t colors, int mono) {
  orc_buffer *b = (orc_buffer *)malloc(sizeof(orc_buffer));
  b->width = w; b->height = h; b->colors = colors; b->monochrome = mono;
  size_t n = w * h * colors;
  b->data = (float *)calloc(n ? n : 1, sizeof(float));
  return b;
}
orc_buffer *orc_buffer_from(size_t w, size_t h, size_t colors, int mono, const float *data) {
  orc_buffer *b = orc_buffer_new(w, h, colors, mono);
  memcpy(b->data, data, w * h * colors * sizeof(float));
  return b;
}
/* src/buffer.rs:45 `self.clone()` — a single-threaded Vec copy */
orc_buffer *orc_buffer_clone(const orc_buffer *s) {
  orc_buffer *b = (orc_buffer *)malloc(sizeof(orc_buffer));
  *b = *s;
  size_t n = s->width * s->height * s->colors;
  b->data = (float *)malloc((n ? n : 1) * sizeof(float));
  memcpy(b->data, s->data, n * sizeof(float));
  return b;
}
void orc_buffer_free(orc_buffer *b) {
  if (!b) return;
  free(b->data);
  free(b);
}
void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------ color_conversions.rs */

/* src/color_conversions.rs:2-6 */
static const float SRGB_D65_33[3][3] = {
  {0.4124564f, 0.3575761f, 0.1804375f},
  {0.2126729f, 0.7151522f, 0.0721750f},
  {0.0193339f, 0.1191920f, 0.9503041f},
};
/* src/color_conversions.rs:7 */
static const float WHITE_X = 0.95047f, WHITE_Y = 1.000f, WHITE_Z = 1.08883f;

/* src/color_conversions.rs:20-39 inverse() in f32 */
static void inverse33(const float m[3][3], float out[3][3]) {
  float invdet = 1.0f / (
    m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) -
    m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
    m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]));
  out[0][0] =  (m[1][1]*m[2][2] - m[2][1]*m[1][2]) * invdet;
  out[0][1] = -(m[0][1]*m[2][2] - m[0][2]*m[2][1]) * invdet;
  out[0][2] =  (m[0][1]*m[1][2] - m[0][2]*m[1][1]) * invdet;
  out[1][0] = -(m[1][0]*m[2][2] - m[1][2]*m[2][0]) * invdet;
  out[1][1] =  (m[0][0]*m[2][2] - m[0][2]*m[2][0]) * invdet;
  out[1][2] = -(m[0][0]*m[1][2] - m[1][0]*m[0][2]) * invdet;
  out[2][0] =  (m[1][0]*m[2][1] - m[2][0]*m[1][1]) * invdet;
  out[2][1] = -(m[0][0]*m[2][1] - m[2][0]*m[0][1]) * invdet;
  out[2][2] =  (m[0][0]*m[1][1] - m[1][0]*m[0][1]) * invdet;
}

static float XYZ_D65_33[3][3];   /* :8  */
static float SRGB_D65_43[3][4];  /* :12-16 */
static float XYZ_D65_34[4][3];   /* :9-11 */

/* src/color_conversions.rs:80-115 TransformLookup */
#define LUT_MAX 8191
static float LUT_XYZ_LAB[LUT_MAX + 2];
static float LUT_SRGB_REV[LUT_MAX + 2];
static float LUT_SRGB_FWD[LUT_MAX + 2];

/* :120-124 */
static float f_xyz_lab(float v) {
  float e = 216.0f / 24389.0f;
  float k = 24389.0f / 27.0f;
  if (v > e) return cbrtf(v);
  return (k * v + 16.0f) / 116.0f;
}
/* :126-132 */
static float f_srgb_reverse(float v) {
  if (v < 0.04045f) return v / 12.92f;
  return powf((v + 0.055f) / 1.055f, 2.4f);
}
/* :134-140 */
static float f_srgb_transform(float v) {
  if (v < 0.0031308f) return v * 12.92f;
  return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}

static int g_init_done = 0;
static void init_statics(void) {
  if (g_init_done) return;
#pragma omp critical(orc_init)
  {
    if (!g_init_done) {
      inverse33(SRGB_D65_33, XYZ_D65_33);
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) { SRGB_D65_43[i][j] = SRGB_D65_33[i][j]; XYZ_D65_34[i][j] = XYZ_D65_33[i][j]; }
        SRGB_D65_43[i][3] = 0.0f;
        XYZ_D65_34[3][i] = 0.0f;
      }
      /* :87-94 table[i] = f(i as f32 / max as f32), i in 0..=max+1 */
      for (int i = 0; i <= LUT_MAX + 1; i++) {
        float v = (float)i / (float)LUT_MAX;
        LUT_XYZ_LAB[i] = f_xyz_lab(v);
        LUT_SRGB_REV[i] = f_srgb_reverse(v);
        LUT_SRGB_FWD[i] = f_srgb_transform(v);
      }
      __sync_synchronize();
      g_init_done = 1;
    }
  }
}
__attribute__((constructor)) static void orc_ctor(void) { init_statics(); }

void orc_matrices(float srgb[9], float xyz[9]) {
  init_statics();
  memcpy(srgb, SRGB_D65_33, sizeof(float) * 9);
  memcpy(xyz, XYZ_D65_33, sizeof(float) * 9);
}
const float *orc_lut_xyz_lab(void) { init_statics(); return LUT_XYZ_LAB; }
const float *orc_lut_srgb_reverse(void) { init_statics(); return LUT_SRGB_REV; }
const float *orc_lut_srgb_transform(void) { init_statics(); return LUT_SRGB_FWD; }

/* :102-114 TransformLookup::lookup */
static inline float lut_lookup(const float *table, float (*f)(float), float val) {
  if (val < 0.0f || val > 1.0f) return f(val);
  float pos = val * (float)LUT_MAX;
  size_t key = f2usize(pos);
  float base = truncf(pos);
  float a = pos - base;
  float v1 = table[key];
  float v2 = table[key + 1];
  return v1 + a * (v2 - v1);
}
/* :144-153 */
float orc_expand_srgb_gamma(float v) { return lut_lookup(LUT_SRGB_REV, f_srgb_reverse, v); }
float orc_apply_srgb_gamma(float v) { return lut_lookup(LUT_SRGB_FWD, f_srgb_transform, v); }

/* :156-169 */
void orc_xyz_to_lab(float x, float y, float z, float out[3]) {
  float xr = x / WHITE_X, yr = y / WHITE_Y, zr = z / WHITE_Z;
  float fx = lut_lookup(LUT_XYZ_LAB, f_xyz_lab, xr);
  float fy = lut_lookup(LUT_XYZ_LAB, f_xyz_lab, yr);
  float fz = lut_lookup(LUT_XYZ_LAB, f_xyz_lab, zr);
  float l = 116.0f * fy - 16.0f;
  float a = 500.0f * (fx - fy);
  float b = 200.0f * (fy - fz);
  out[0] = l / 100.0f;
  out[1] = (a + 127.0f) / 255.0f;
  out[2] = (b + 127.0f) / 255.0f;
}

/* :172-191 */
void orc_lab_to_xyz(float l, float a, float b, float out[3]) {
  float cl = l * 100.0f;
  float ca = (a * 255.0f) - 127.0f;
  float cb = (b * 255.0f) - 127.0f;
  float fy = (cl + 16.0f) / 116.0f;
  float fx = ca / 500.0f + fy;
  float fz = fy - (cb / 200.0f);
  float e = 216.0f / 24389.0f;
  float k = 24389.0f / 27.0f;
  float fx3 = fx * fx * fx;
  float xr = fx3 > e ? fx3 : (116.0f * fx - 16.0f) / k;
  float yr = cl > k * e ? fy * fy * fy : cl / k;
  float fz3 = fz * fz * fz;
  float zr = fz3 > e ? fz3 : (116.0f * fz - 16.0f) / k;
  out[0] = xr * WHITE_X;
  out[1] = yr * WHITE_Y;
  out[2] = zr * WHITE_Z;
}

/* :42-55 — cmatrix is [[f32;4];3] row-major */
void orc_camera_to_lab(const float mul[4], const float cm[12], const float pixin[4], float out[3]) {
  float r = fminf(pixin[0] * mul[0], 1.0f);
  float g = fminf(pixin[1] * mul[1], 1.0f);
  float b = fminf(pixin[2] * mul[2], 1.0f);
  float e = fminf(pixin[3] * mul[3], 1.0f);
  float x = r * cm[0] + g * cm[1] + b * cm[2] + e * cm[3];
  float y = r * cm[4] + g * cm[5] + b * cm[6] + e * cm[7];
  float z = r * cm[8] + g * cm[9] + b * cm[10] + e * cm[11];
  orc_xyz_to_lab(x, y, z, out);
}

/* :58-65 — rgbmatrix is [[f32;3];3] row-major */
void orc_lab_to_rgb(const float m[9], const float pixin[3], float out[3]) {
  float xyz[3];
  orc_lab_to_xyz(pixin[0], pixin[1], pixin[2], xyz);
  float x = xyz[0], y = xyz[1], z = xyz[2];
  out[0] = x * m[0] + y * m[1] + z * m[2];
  out[1] = x * m[3] + y * m[4] + z * m[5];
  out[2] = x * m[6] + y * m[7] + z * m[8];
}

/* :312-330 */
float orc_input8bit(uint8_t v) { return (float)v / 255.0f; }
float orc_input16bit(uint16_t v) { return (float)v / 65535.0f; }
uint8_t orc_output8bit(float v) {
  /* Rust f32::max/min ignore NaN like fmaxf/fminf; `as u8` truncates */
  return (uint8_t)fminf(fmaxf(v * 256.0f, 0.0f), 255.0f);
}
uint16_t orc_output16bit(float v) {
  return (uint16_t)fminf(fmaxf(roundf(v * 65535.0f), 0.0f), 65535.0f);
}

/* ------------------------------------------------------------------ rawloader::CFA
 * rawloader 0.37 is a path dependency that is NOT in the tree (Cargo.toml:25-27).
 * Call sites: demosaic.rs:32-33,80,86; scaling.rs:110.  Published behaviour restated:
 * pattern length 0/4/36/16/144 -> 0x0 / 2x2 / 6x6 / 2 wide x 8 high / 12x12,
 * chars R,G,B,E -> 0,1,2,3 (M -> 1, Y -> 3), tiled into 48x48;
 * color_at(row,col) = pattern[row % 48][col % 48].  PARITY UNPINNED by reference tests. */
int orc_cfa_new(orc_cfa *cfa, const char *pat) {
  size_t len = strlen(pat);
  size_t w, h;
  switch (len) {
    case 0: w = 0; h = 0; break;
    case 4: w = 2; h = 2; break;
    case 36: w = 6; h = 6; break;
    case 16: w = 2; h = 8; break;
    case 144: w = 12; h = 12; break;
    default: return -1;
  }
  memset(cfa, 0, sizeof(*cfa));
  cfa->width = w; cfa->height = h;
  if (w > 0) {
    for (size_t i = 0; i < len; i++) {
      uint8_t c;
      switch (pat[i]) {
        case 'R': c = 0; break;
        case 'G': c = 1; break;
        case 'B': c = 2; break;
        case 'E': c = 3; break;
        case 'M': c = 1; break;
        case 'Y': c = 3; break;
        default: return -2;
      }
      cfa->pattern[i / w][i % w] = c;
    }
    for (size_t r = 0; r < 48; r++)
      for (size_t c = 0; c < 48; c++)
        cfa->pattern[r][c] = cfa->pattern[r % h][c % w];
  }
  return 0;
}
size_t orc_cfa_color_at(const orc_cfa *cfa, size_t row, size_t col) {
  return cfa->pattern[(row + 48) % 48][(col + 48) % 48];
}

/* ------------------------------------------------------------------ scaling.rs */

/* src/scaling.rs:8-23 */
static void calculate_scaling_total(size_t width, size_t height, size_t maxwidth, size_t maxheight,
                                    float *scale, size_t *ow, size_t *oh) {
  if (maxwidth == 0 && maxheight == 0) { *scale = 1.0f; *ow = width; *oh = height; return; }
  float xscale = maxwidth == 0 ? 1.0f : (float)width / (float)maxwidth;
  float yscale = maxheight == 0 ? 1.0f : (float)height / (float)maxheight;
  if (yscale <= 1.0f && xscale <= 1.0f) { *scale = 1.0f; *ow = width; *oh = height; }
  else if (yscale > xscale) { *scale = yscale; *ow = f2usize((float)width / yscale); *oh = maxheight; }
  else { *scale = xscale; *ow = maxwidth; *oh = f2usize((float)height / xscale); }
}
/* :25-28 */
void orc_scaling_size(size_t w, size_t h, size_t maxw, size_t maxh, size_t *ow, size_t *oh) {
  float s; calculate_scaling_total(w, h, maxw, maxh, &s, ow, oh);
}
/* :30-32 */
float orc_calculate_scale(size_t w, size_t h, size_t maxw, size_t maxh) {
  float s; size_t a, b; calculate_scaling_total(w, h, maxw, maxh, &s, &a, &b); return s;
}

/* src/scaling.rs:51-130 transform_buffer<T>.  AS_F32 = T::as_() to f32, FROM_F32 = f32::as_() to T. */
#define DEFINE_TRANSFORM_BUFFER(NAME, T, AS_F32, FROM_F32)                                             \
  static void NAME(const T *src, size_t width, size_t height, const long tl[2], const long tr[2],      \
                   const long bl[2], size_t nwidth, size_t nheight, size_t components,                 \
                   const orc_cfa *cfa, T *out) {                                                       \
    memset(out, 0, nwidth * nheight * components * sizeof(T));                                         \
    const float tl0 = (float)tl[0], tl1 = (float)tl[1];                                                \
    const float skip_x_x = ((float)tr[0] - tl0) / (float)(nwidth - 1);                                 \
    const float skip_x_y = ((float)tr[1] - tl1) / (float)(nwidth - 1);                                 \
    const float skip_y_x = ((float)bl[0] - tl0) / (float)(nheight - 1);                                \
    const float skip_y_y = ((float)bl[1] - tl1) / (float)(nheight - 1);                                \
    _Pragma("omp parallel for schedule(dynamic, 8)")                                                   \
    for (size_t row = 0; row < nheight; row++) {                                                       \
      T *line = out + row * nwidth * components;                                                       \
      const float rfrom_x = tl0 + skip_y_x * (float)row;                                               \
      const float rto_x = tl0 + skip_y_x * (float)(row + 1);                                           \
      const float rfrom_y = tl1 + skip_y_y * (float)row;                                               \
      const float rto_y = tl1 + skip_y_y * (float)(row + 1);                                           \
      const float rcenter_x = tl0 + (skip_y_x * (float)row) + (skip_y_x / 2.0f) - 0.5f;                \
      const float rcenter_y = tl1 + (skip_y_y * (float)row) + (skip_y_y / 2.0f) - 0.5f;                \
      for (size_t col = 0; col < nwidth; col++) {                                                      \
        size_t from_x = umin(width - 1, f2usize(floorf(rfrom_x + (skip_x_x * (float)col))));           \
        size_t to_x = umin(width - 1, f2usize(floorf(rto_x + (skip_x_x * (float)(col + 1)))));         \
        size_t from_y = umin(height - 1, f2usize(floorf(rfrom_y + (skip_x_y * (float)col))));          \
        size_t to_y = umin(height - 1, f2usize(floorf(rto_y + (skip_x_y * (float)(col + 1)))));        \
        float center_x = rcenter_x + (skip_x_x * (float)col) + (skip_x_x / 2.0f);                      \
        float center_y = rcenter_y + (skip_x_y * (float)col) + (skip_x_y / 2.0f);                      \
        float sums[4] = {0.0f, 0.0f, 0.0f, 0.0f};                                                      \
        float counts[4] = {0.0f, 0.0f, 0.0f, 0.0f};                                                    \
        for (size_t y = from_y; y <= to_y; y++) {                                                      \
          for (size_t x = from_x; x <= to_x; x++) {                                                    \
            float delta_x = ((float)x - center_x) / skip_x_x;                                          \
            float delta_y = ((float)y - center_y) / skip_y_y;                                          \
            float factor = 1.0f - (delta_x * delta_x) - (delta_y * delta_y);                           \
            factor = factor < 0.0f ? 0.0f : factor;                                                    \
            if (cfa) {                                                                                 \
              size_t c = orc_cfa_color_at(cfa, y, x);                                                  \
              sums[c] += AS_F32(src[y * width + x]) * factor;                                          \
              counts[c] += factor;                                                                     \
            } else {                                                                                   \
              for (size_t c = 0; c < components; c++) {                                                \
                sums[c] += AS_F32(src[(y * width + x) * components + c]) * factor;                     \
                counts[c] += factor;                                                                   \
              }                                                                                        \
            }                                                                                          \
          }                                                                                            \
        }                                                                                              \
        for (size_t c = 0; c < components; c++)                                                        \
          if (counts[c] > 0.0f) line[col * components + c] = FROM_F32(sums[c] / counts[c]);            \
      }                                                                                                \
    }                                                                                                  \
  }

static inline float id_f32(float v) { return v; }
static inline float u8_as_f32(uint8_t v) { return (float)v; }
static inline float u16_as_f32(uint16_t v) { return (float)v; }
/* Rust `f32 as u8/u16` saturates, NaN -> 0 */
static inline uint8_t f32_as_u8(float v) { return v != v ? 0 : v <= 0.0f ? 0 : v >= 255.0f ? 255 : (uint8_t)v; }
static inline uint16_t f32_as_u16(float v) { return v != v ? 0 : v <= 0.0f ? 0 : v >= 65535.0f ? 65535 : (uint16_t)v; }

DEFINE_TRANSFORM_BUFFER(transform_buffer_f32, float, id_f32, id_f32)
DEFINE_TRANSFORM_BUFFER(transform_buffer_u8, uint8_t, u8_as_f32, f32_as_u8)
DEFINE_TRANSFORM_BUFFER(transform_buffer_u16, uint16_t, u16_as_f32, f32_as_u16)

void orc_transform_buffer_f32(const float *src, size_t width, size_t height, const long tl[2],
                              const long tr[2], const long bl[2], size_t nw, size_t nh, size_t comps,
                              const orc_cfa *cfa, float *out) {
  transform_buffer_f32(src, width, height, tl, tr, bl, nw, nh, comps, cfa, out);
}
void orc_transform_buffer_u8(const uint8_t *src, size_t width, size_t height, const long tl[2],
                             const long tr[2], const long bl[2], size_t nw, size_t nh, size_t comps, uint8_t *out) {
  transform_buffer_u8(src, width, height, tl, tr, bl, nw, nh, comps, NULL, out);
}
void orc_transform_buffer_u16(const uint16_t *src, size_t width, size_t height, const long tl[2],
                              const long tr[2], const long bl[2], size_t nw, size_t nh, size_t comps, uint16_t *out) {
  transform_buffer_u16(src, width, height, tl, tr, bl, nw, nh, comps, NULL, out);
}

/* src/scaling.rs:35-48 scale_down_buffer corner points */
#define SCALE_DOWN_CORNERS(w, h) \
  const long tl[2] = {0, 0}, tr[2] = {(long)(w) - 1, 0}, bl[2] = {0, (long)(h) - 1}

/* src/scaling.rs:132-145 */
orc_buffer *orc_scaled_demosaic(const orc_cfa *cfa, const orc_buffer *buf, size_t nw, size_t nh) {
  if (buf->colors != 1) return NULL; /* assert_eq!(buf.colors, 1) */
  orc_buffer *out = orc_buffer_new(nw, nh, 4, buf->monochrome);
  SCALE_DOWN_CORNERS(buf->width, buf->height);
  transform_buffer_f32(buf->data, buf->width, buf->height, tl, tr, bl, nw, nh, 4, cfa, out->data);
  return out;
}
/* src/scaling.rs:147-160 */
orc_buffer *orc_scale_down_opbuf(const orc_buffer *buf, size_t nw, size_t nh) {
  if (buf->colors != 4) return NULL; /* assert_eq!(buf.colors, 4) */
  orc_buffer *out = orc_buffer_new(nw, nh, 4, buf->monochrome);
  SCALE_DOWN_CORNERS(buf->width, buf->height);
  transform_buffer_f32(buf->data, buf->width, buf->height, tl, tr, bl, nw, nh, 4, NULL, out->data);
  return out;
}
/* src/scaling.rs:162-171 */
void orc_scale_down_srgb(const uint8_t *src, size_t w, size_t h, size_t nw, size_t nh, uint8_t *out) {
  SCALE_DOWN_CORNERS(w, h);
  transform_buffer_u8(src, w, h, tl, tr, bl, nw, nh, 3, NULL, out);
}
/* src/scaling.rs:173-182 */
void orc_scale_down_srgb16(const uint16_t *src, size_t w, size_t h, size_t nw, size_t nh, uint16_t *out) {
  SCALE_DOWN_CORNERS(w, h);
  transform_buffer_u16(src, w, h, tl, tr, bl, nw, nh, 3, NULL, out);
}

/* ------------------------------------------------------------------ ops/gofloat.rs */

/* src/ops/gofloat.rs:74-82 */
void orc_gofloat_size_image(const orc_gofloat *op, size_t ow, size_t oh, size_t o[4]) {
  o[0] = umin(op->crop_left, ow - 10);
  o[1] = umin(op->crop_top, oh - 10);
  o[2] = ow - umin(op->crop_left + op->crop_right, ow - 10);
  o[3] = oh - umin(op->crop_top + op->crop_bottom, oh - 10);
}

#define GOFLOAT_RAW_BODY(T)                                                                             \
  const T *data = (const T *)img->data;                                                                 \
  const size_t total = img->width * img->height * img->cpp;                                             \
  if (img->cpp == 1 && !op->is_cfa) { /* :97-109 / :134-145 monochrome -> RGB */                        \
    out = orc_buffer_new(width, height, 4, 1);                                                          \
    _Pragma("omp parallel for schedule(dynamic, 16)")                                                   \
    for (size_t row = 0; row < height; row++) {                                                         \
      float *line = out->data + row * width * 4;                                                        \
      const T *in = data + owidth * (row + y) + x;                                                      \
      for (size_t c = 0; c < width; c++) {                                                              \
        float val = fminf(((float)in[c] - mins[0]) / ranges[0], 1.0f);                                  \
        line[c * 4 + 0] = val; line[c * 4 + 1] = val; line[c * 4 + 2] = val; line[c * 4 + 3] = 0.0f;    \
      }                                                                                                 \
    }                                                                                                   \
  } else if (img->cpp == 3) { /* :110-121 / :146-157 RGB -> four channel */                             \
    out = orc_buffer_new(width, height, 4, 0);                                                          \
    _Pragma("omp parallel for schedule(dynamic, 16)")                                                   \
    for (size_t row = 0; row < height; row++) {                                                         \
      float *line = out->data + row * width * 4;                                                        \
      const T *in = data + (owidth * (row + y) + x) * 3;                                                \
      for (size_t c = 0; c < width; c++) {                                                              \
        line[c * 4 + 0] = fminf(((float)in[c * 3 + 0] - mins[0]) / ranges[0], 1.0f);                    \
        line[c * 4 + 1] = fminf(((float)in[c * 3 + 1] - mins[1]) / ranges[1], 1.0f);                    \
        line[c * 4 + 2] = fminf(((float)in[c * 3 + 2] - mins[2]) / ranges[2], 1.0f);                    \
        line[c * 4 + 3] = 0.0f;                                                                         \
      }                                                                                                 \
    }                                                                                                   \
  } else { /* :122-130 / :158-166 CFA (or anything else): cpp channels, only level index 0 */           \
    out = orc_buffer_new(width, height, img->cpp, 0);                                                   \
    const size_t linelen = width * img->cpp;                                                            \
    _Pragma("omp parallel for schedule(dynamic, 16)")                                                   \
    for (size_t row = 0; row < height; row++) {                                                         \
      float *line = out->data + row * linelen;                                                          \
      const size_t off = owidth * (row + y) + x;                                                        \
      /* zip() stops at the shorter of the two iterators */                                             \
      size_t n = off < total ? umin(linelen, total - off) : 0;                                          \
      for (size_t c = 0; c < n; c++)                                                                    \
        line[c] = fminf(((float)data[off + c] - mins[0]) / ranges[0], 1.0f);                            \
    }                                                                                                   \
  }

/* src/ops/gofloat.rs:84-169 run_raw, :171-201 run_other, :50-62 run */
orc_buffer *orc_gofloat_run(const orc_gofloat *op, const orc_source *img) {
  init_statics();
  const size_t owidth = img->width, oheight = img->height;
  size_t xywh[4];
  orc_gofloat_size_image(op, owidth, oheight, xywh);
  const size_t x = xywh[0], y = xywh[1], width = xywh[2], height = xywh[3];
  orc_buffer *out = NULL;
  if (img->kind == ORC_SRC_RAW_U16 || img->kind == ORC_SRC_RAW_F32) {
    float mins[4], ranges[4];
    for (int i = 0; i < 4; i++) { mins[i] = op->blacklevels[i]; ranges[i] = op->whitelevels[i] - mins[i]; }
    if (img->kind == ORC_SRC_RAW_U16) { GOFLOAT_RAW_BODY(uint16_t) }
    else { GOFLOAT_RAW_BODY(float) }
  } else if (img->kind == ORC_SRC_RGB8) { /* :177-186 */
    const uint8_t *data = (const uint8_t *)img->data;
    out = orc_buffer_new(width, height, 4, 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t row = 0; row < height; row++) {
      float *line = out->data + row * width * 4;
      const uint8_t *in = data + (owidth * (row + y) + x) * 3;
      for (size_t c = 0; c < width; c++) {
        line[c * 4 + 0] = orc_expand_srgb_gamma(orc_input8bit(in[c * 3 + 0]));
        line[c * 4 + 1] = orc_expand_srgb_gamma(orc_input8bit(in[c * 3 + 1]));
        line[c * 4 + 2] = orc_expand_srgb_gamma(orc_input8bit(in[c * 3 + 2]));
        line[c * 4 + 3] = 0.0f;
      }
    }
  } else { /* :187-197 */
    const uint16_t *data = (const uint16_t *)img->data;
    out = orc_buffer_new(width, height, 4, 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t row = 0; row < height; row++) {
      float *line = out->data + row * width * 4;
      const uint16_t *in = data + (owidth * (row + y) + x) * 3;
      for (size_t c = 0; c < width; c++) {
        line[c * 4 + 0] = orc_input16bit(in[c * 3 + 0]);
        line[c * 4 + 1] = orc_input16bit(in[c * 3 + 1]);
        line[c * 4 + 2] = orc_input16bit(in[c * 3 + 2]);
        line[c * 4 + 3] = 0.0f;
      }
    }
  }
  return out;
}

/* ------------------------------------------------------------------ ops/demosaic.rs */

/* src/ops/demosaic.rs:67-119 full() */
orc_buffer *orc_demosaic_full(const orc_cfa *cfa, const orc_buffer *buf) {
  orc_buffer *out = orc_buffer_new(buf->width, buf->height, 4, buf->monochrome);
  static const int off[9][2] = { /* (dy,dx) :70-74 */
    {-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 0}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
  /* :77-90 colour of each 3x3 tap; same colour as the centre (but not the centre) -> bin 4 */
  static __thread uint8_t lookups[48][48][9];
  for (size_t row = 0; row < 48; row++)
    for (size_t col = 0; col < 48; col++) {
      size_t pixcolor = orc_cfa_color_at(cfa, row, col);
      for (int i = 0; i < 9; i++) {
        int dy = off[i][0], dx = off[i][1];
        size_t ocolor = orc_cfa_color_at(cfa, (size_t)(48 + dy) + row, (size_t)(48 + dx) + col);
        lookups[row][col][i] = (ocolor != pixcolor || (dx == 0 && dy == 0)) ? (uint8_t)ocolor : 4;
      }
    }
  const uint8_t(*lk)[48][9] = lookups;
  const long h = (long)buf->height, w = (long)buf->width;
  /* :93-116 */
#pragma omp parallel for schedule(dynamic, 16)
  for (long row = 0; row < h; row++) {
    float *line = out->data + (size_t)row * buf->width * 4;
    for (long col = 0; col < w; col++) {
      const uint8_t *colors = lk[row % 48][col % 48];
      float sums[5] = {0, 0, 0, 0, 0}, counts[5] = {0, 0, 0, 0, 0};
      for (int i = 0; i < 9; i++) {
        long r = row + off[i][0], c = col + off[i][1];
        if (r >= 0 && r < h && c >= 0 && c < w) {
          sums[colors[i]] += buf->data[(size_t)r * buf->width + (size_t)c];
          counts[colors[i]] += 1.0f;
        }
      }
      for (int c = 0; c < 4; c++)
        if (counts[c] > 0.0f) line[col * 4 + c] = sums[c] / counts[c];
    }
  }
  return out;
}

/* src/ops/demosaic.rs:27-61 run */
orc_buffer *orc_demosaic_run(const orc_demosaic *op, const orc_settings *s, orc_buffer *buf) {
  size_t nwidth = s->demosaic_width, nheight = s->demosaic_height;
  float scale = orc_calculate_scale(buf->width, buf->height, nwidth, nheight);
  orc_cfa cfa;
  if (orc_cfa_new(&cfa, op->cfa) != 0) return NULL;
  float minscale;
  switch (cfa.width) {
    case 2: minscale = 2.0f; break;
    case 6: minscale = 3.0f; break;
    case 8: minscale = 2.0f; break;
    case 12: minscale = 12.0f; break;
    default: minscale = 2.0f; break;
  }
  if (scale <= 1.0f && buf->colors == 4) return buf;
  if (buf->colors == 4) return orc_scale_down_opbuf(buf, nwidth, nheight);
  if (scale >= minscale) return orc_scaled_demosaic(&cfa, buf, nwidth, nheight);
  orc_buffer *fullsize = orc_demosaic_full(&cfa, buf);
  if (scale > 1.0f) {
    orc_buffer *o = orc_scale_down_opbuf(fullsize, nwidth, nheight);
    orc_buffer_free(fullsize);
    return o;
  }
  return fullsize;
}

/* ------------------------------------------------------------------ ops/rotatecrop.rs */

static const float RC_EPSILON = 1.0f / 1000000.0f; /* :7 */
#define FRAC_PI_2 1.57079632679489661923132169163975144f

/* :89-95 */
static int rc_noop(const orc_rotatecrop *op) {
  return fabsf(op->rotation) < RC_EPSILON && fabsf(op->crop_top) < RC_EPSILON &&
         fabsf(op->crop_right) < RC_EPSILON && fabsf(op->crop_bottom) < RC_EPSILON &&
         fabsf(op->crop_left) < RC_EPSILON;
}
/* :97-109 */
static void rc_rotate_point_reverse(const orc_rotatecrop *op, float x, float y, float width, float height,
                                    float swidth, float sheight, long out[2]) {
  if (op->rotation < RC_EPSILON) { out[0] = f2isize(x); out[1] = f2isize(y); return; }
  float angle = FRAC_PI_2 * (op->rotation > 1.0f ? 1.0f : op->rotation);
  float sn = sinf(angle), cs = cosf(angle);
  float tx = x - (width / 2.0f), ty = y - (height / 2.0f);
  float nx = tx * cs + ty * sn + (swidth / 2.0f);
  float ny = -tx * sn + ty * cs + (sheight / 2.0f);
  out[0] = f2isize(nx); out[1] = f2isize(ny);
}
/* :111-163 */
static void rc_calc_size(const orc_rotatecrop *op, size_t owidth, size_t oheight, int reverse, size_t *ow, size_t *oh) {
  if (rc_noop(op)) { *ow = owidth; *oh = oheight; return; }
  float width = (float)owidth, height = (float)oheight;
  if (!(reverse || op->rotation < RC_EPSILON)) {
    float angle = FRAC_PI_2 * (op->rotation > 1.0f ? 1.0f : op->rotation);
    float sn = sinf(angle), cs = cosf(angle);
    float w2 = width * cs + height * sn, h2 = width * sn + height * cs;
    width = w2; height = h2;
  }
  float nwidth, nheight;
  {
    float ratio = 1.0f - op->crop_left - op->crop_right;
    nwidth = reverse ? roundf(width / ratio) : roundf(width * ratio);
    if (ratio < RC_EPSILON || nwidth < 1.0f) { *ow = owidth; *oh = oheight; return; }
  }
  {
    float ratio = 1.0f - op->crop_top - op->crop_bottom;
    nheight = reverse ? roundf(height / ratio) : roundf(height * ratio);
    if (ratio < RC_EPSILON || nheight < 1.0f) { *ow = owidth; *oh = oheight; return; }
  }
  if (!(!reverse || op->rotation < RC_EPSILON)) {
    float angle = FRAC_PI_2 * (op->rotation > 1.0f ? 1.0f : op->rotation);
    float sn = sinf(angle), cs = cosf(angle);
    float w2 = roundf(nheight / (sn + (cs / op->input_ratio)));
    float h2 = roundf(w2 / op->input_ratio);
    nwidth = w2; nheight = h2;
  }
  *ow = f2usize(nwidth); *oh = f2usize(nheight);
}
/* :66-74 */
void orc_rotatecrop_transform_forward(orc_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh) {
  if (op->has_output_size) { *ow = op->output_width; *oh = op->output_height; return; }
  op->input_ratio = (float)w / (float)h;
  rc_calc_size(op, w, h, 0, ow, oh);
}
/* :76-80 */
void orc_rotatecrop_transform_reverse(orc_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh) {
  op->has_output_size = 1; op->output_width = w; op->output_height = h;
  rc_calc_size(op, w, h, 1, ow, oh);
}
/* :82-85 */
void orc_rotatecrop_reset(orc_rotatecrop *op) { op->input_ratio = 1.0f; op->has_output_size = 0; }

/* :39-64 run; buffer.rs:62-79 OpBuffer::transform */
orc_buffer *orc_rotatecrop_run(const orc_rotatecrop *op, orc_buffer *buf) {
  if (rc_noop(op)) return buf;
  float swidth = (float)buf->width, sheight = (float)buf->height;
  size_t nwidth, nheight;
  rc_calc_size(op, buf->width, buf->height, 0, &nwidth, &nheight);
  float fnwidth = (float)nwidth, fnheight = (float)nheight;
  float x = floorf(swidth * op->crop_left);
  if (x < 0.0f || x > swidth) return buf;
  float y = floorf(sheight * op->crop_top);
  if (y < 0.0f || y > sheight) return buf;
  long tl[2], tr[2], bl[2];
  rc_rotate_point_reverse(op, x, y, fnwidth, fnheight, swidth, sheight, tl);
  rc_rotate_point_reverse(op, x + fnwidth - 1.0f, y, fnwidth, fnheight, swidth, sheight, tr);
  rc_rotate_point_reverse(op, x, y + fnheight - 1.0f, fnwidth, fnheight, swidth, sheight, bl);
  orc_buffer *out = orc_buffer_new(nwidth, nheight, buf->colors, buf->monochrome);
  transform_buffer_f32(buf->data, buf->width, buf->height, tl, tr, bl, nwidth, nheight, buf->colors, NULL, out->data);
  return out;
}

/* ------------------------------------------------------------------ ops/colorspaces.rs */

/* :12-27 */
static void normalize_wbs(const float vals[4], float out[4]) {
  float unity = vals[1];
  for (int i = 0; i < 4; i++) out[i] = !isnormal(vals[i]) ? 1.0f : vals[i] / unity;
}
/* :89-112 OpToLab::run (buffer.rs:52-60 process_into_new) */
orc_buffer *orc_tolab_run(const orc_tolab *op, orc_buffer *buf) {
  init_statics();
  float cm[12], mul[4];
  if (buf->monochrome) {
    memcpy(cm, SRGB_D65_43, sizeof(cm));
    mul[0] = mul[1] = mul[2] = mul[3] = 1.0f;
  } else {
    memcpy(cm, op->cam_to_xyz_normalized, sizeof(cm));
    normalize_wbs(op->wb_coeffs, mul);
  }
  orc_buffer *out = orc_buffer_new(buf->width, buf->height, 3, buf->monochrome);
  const size_t w = buf->width, h = buf->height;
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t row = 0; row < h; row++) {
    const float *in = buf->data + row * w * 4;
    float *o = out->data + row * w * 3;
    for (size_t c = 0; c < w; c++) orc_camera_to_lab(mul, cm, in + c * 4, o + c * 3);
  }
  return out;
}
/* :127-137 OpFromLab::run (buffer.rs:42-50 mutate_lines_copying: clone, then mutate) */
orc_buffer *orc_fromlab_run(orc_buffer *buf) {
  init_statics();
  orc_buffer *out = orc_buffer_clone(buf);
  const size_t w = buf->width, h = buf->height;
  const float *m = &XYZ_D65_33[0][0];
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t row = 0; row < h; row++) {
    float *line = out->data + row * w * 3;
    for (size_t c = 0; c < w; c++) {
      float rgb[3];
      orc_lab_to_rgb(m, line + c * 3, rgb);
      line[c * 3 + 0] = rgb[0]; line[c * 3 + 1] = rgb[1]; line[c * 3 + 2] = rgb[2];
    }
  }
  return out;
}

/* ------------------------------------------------------------------ ops/curves.rs */

/* :68-124 SplineFunc::new */
void orc_spline_new(orc_spline *s, const float (*p)[2], size_t n) {
  memset(s, 0, sizeof(*s));
  size_t np = 0;
  if (n == 0 || (p[0][0] > 0.0f && p[0][1] > 0.0f)) { s->x[np] = 0.0f; s->y[np] = 0.0f; np++; }
  for (size_t i = 0; i < n; i++) { s->x[np] = p[i][0]; s->y[np] = p[i][1]; np++; }
  if (n == 0 || (p[n - 1][0] < 1.0f && p[n - 1][1] < 1.0f)) { s->x[np] = 1.0f; s->y[np] = 1.0f; np++; }
  s->n = np;
  if (np < 2) { s->nseg = 0; return; } /* the reference would panic on slopes[0] */
  float dxs[ORC_MAX_CURVE_POINTS + 2], slopes[ORC_MAX_CURVE_POINTS + 2];
  size_t nd = np - 1;
  for (size_t i = 0; i < nd; i++) {
    float dx = s->x[i + 1] - s->x[i];
    float dy = s->y[i + 1] - s->y[i];
    dxs[i] = dx;
    slopes[i] = dy / dx;
  }
  size_t nc1 = 0;
  s->c1[nc1++] = slopes[0];
  for (size_t i = 0; i + 1 < nd; i++) {
    float m = slopes[i], next = slopes[i + 1];
    if (m * next <= 0.0f) s->c1[nc1++] = 0.0f;
    else {
      float dx = dxs[i], dxnext = dxs[i + 1];
      float common = dx + dxnext;
      s->c1[nc1++] = 3.0f * common / ((common + dxnext) / m + (common + dx) / next);
    }
  }
  s->c1[nc1++] = slopes[nd - 1];
  for (size_t i = 0; i + 1 < nc1; i++) {
    float c1 = s->c1[i], slope = slopes[i];
    float invdx = 1.0f / dxs[i];
    float common = c1 + s->c1[i + 1] - slope - slope;
    s->c2[i] = (slope - c1 - common) * invdx;
    s->c3[i] = common * invdx * invdx;
  }
  s->nseg = nc1 - 1;
}
/* :126-157 interpolate */
float orc_spline_interpolate(const orc_spline *s, float val) {
  float end = s->x[s->n - 1];
  if (val >= end) return s->y[s->n - 1];
  float first = s->x[0];
  if (val <= first) return s->y[0];
  long low = 0, mid, high = (long)s->nseg - 1;
  while (low <= high) {
    mid = (low + high) / 2;
    float xhere = s->x[mid];
    if (xhere < val) low = mid + 1;
    else if (xhere > val) high = mid - 1;
    else return s->y[mid];
  }
  size_t i = (size_t)(high > 0 ? high : 0);
  float diff = val - s->x[i];
  return s->y[i] + s->c1[i] * diff + s->c2[i] * diff * diff + s->c3[i] * diff * diff * diff;
}
/* :33-49 OpBaseCurve::run */
orc_buffer *orc_basecurve_run(const orc_basecurve *op, orc_buffer *buf) {
  if (op->npoints == 0 && fabsf(op->exposure) < 0.001f) return buf;
  float pts[ORC_MAX_CURVE_POINTS][2];
  float ex = exp2f(op->exposure);
  for (size_t i = 0; i < op->npoints; i++) { pts[i][0] = op->points[i][0]; pts[i][1] = op->points[i][1] * ex; }
  orc_spline sp;
  orc_spline_new(&sp, (const float(*)[2])pts, op->npoints);
  orc_buffer *out = orc_buffer_clone(buf);
  const size_t w = buf->width, h = buf->height;
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t row = 0; row < h; row++) {
    float *line = out->data + row * w * 3;
    for (size_t c = 0; c < w; c++) line[c * 3] = orc_spline_interpolate(&sp, line[c * 3]);
  }
  return out;
}

/* ------------------------------------------------------------------ ops/gamma.rs */

/* :16-26 */
orc_buffer *orc_gamma_run(const orc_settings *s, orc_buffer *buf) {
  if (s->linear) return buf;
  init_statics();
  orc_buffer *out = orc_buffer_clone(buf);
  const size_t n = buf->width * buf->colors, h = buf->height;
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t row = 0; row < h; row++) {
    float *line = out->data + row * n;
    for (size_t c = 0; c < n; c++) line[c] = orc_apply_srgb_gamma(fminf(fmaxf(line[c], 0.0f), 1.0f));
  }
  return out;
}

/* ------------------------------------------------------------------ ops/transform.rs */

/* rawloader Orientation::to_flips -> (transpose, flip_x, flip_y), pinned by the eight
 * golden bitmaps of transform.rs:167-278 (see tests/test_oracle_kats.py). */
static void rotation_to_flips(int rotation, int f[3]) {
  switch (rotation) {
    case ORC_ROT_90: f[0] = 1; f[1] = 0; f[2] = 1; break;
    case ORC_ROT_180: f[0] = 0; f[1] = 1; f[2] = 1; break;
    case ORC_ROT_270: f[0] = 1; f[1] = 1; f[2] = 0; break;
    default: f[0] = 0; f[1] = 0; f[2] = 0; break;
  }
}
/* transform.rs:56-66: base orientation flips XOR user flips */
void orc_orientation_flips(const orc_transform *op, int f[3]) {
  rotation_to_flips(op->rotation, f);
  f[1] ^= (op->fliph != 0);
  f[2] ^= (op->flipv != 0);
}
/* transform.rs:87-144 rotate_buffer */
orc_buffer *orc_rotate_buffer(const orc_buffer *buf, int transpose, int flip_x, int flip_y) {
  if (buf->colors != 3) return NULL; /* assert_eq!(buf.colors, 3) */
  if (!transpose && !flip_x && !flip_y) return orc_buffer_clone(buf);
  long width = (long)buf->width, height = (long)buf->height;
  long base_offset = 0, x_step = 3, y_step = width * 3;
  if (flip_x) { x_step = -x_step; base_offset += (width - 1) * 3; }
  if (flip_y) { y_step = -y_step; base_offset += width * (height - 1) * 3; }
  orc_buffer *out;
  if (transpose) {
    long t = width; width = height; height = t;
    t = x_step; x_step = y_step; y_step = t;
    out = orc_buffer_new(buf->height, buf->width, 3, buf->monochrome);
  } else {
    out = orc_buffer_new(buf->width, buf->height, 3, buf->monochrome);
  }
#pragma omp parallel for schedule(dynamic, 16)
  for (long row = 0; row < height; row++) {
    float *line = out->data + (size_t)row * (size_t)width * 3;
    long line_offset = base_offset + y_step * row;
    for (long col = 0; col < width; col++) {
      long offset = line_offset + x_step * col;
      for (long c = 0; c < 3; c++) line[col * 3 + c] = buf->data[offset + c];
    }
  }
  return out;
}
/* transform.rs:56-73 run */
orc_buffer *orc_transform_run(const orc_transform *op, orc_buffer *buf) {
  int f[3];
  orc_orientation_flips(op, f);
  if (!f[0] && !f[1] && !f[2]) return buf;
  return orc_rotate_buffer(buf, f[0], f[1], f[2]);
}
/* transform.rs:75-84 */
void orc_transform_transform_forward(const orc_transform *op, size_t w, size_t h, size_t *ow, size_t *oh) {
  if (op->rotation == ORC_ROT_90 || op->rotation == ORC_ROT_270) { *ow = h; *oh = w; }
  else { *ow = w; *oh = h; }
}

/* ------------------------------------------------------------------ pipeline.rs */

/* PipelineOps::new (pipeline.rs:166-179) for the parts that do not need rawloader
 * metadata: gofloat.rs:20-47, demosaic.rs:9-23, rotatecrop.rs:24-34, colorspaces.rs:48-55,
 * curves.rs:12-28, transform.rs:43-50.  Raw metadata (levels, cfa, matrices, wb) is
 * filled in by the caller as literals. */
void orc_pipeline_defaults(orc_pipeline *p, const orc_source *img) {
  init_statics();
  memset(p, 0, sizeof(*p));
  p->image = *img;
  p->settings.use_fastpath = 1; /* pipeline.rs:121-130 */
  p->ops.rotatecrop.input_ratio = 1.0f;
  int raw = img->kind == ORC_SRC_RAW_U16 || img->kind == ORC_SRC_RAW_F32;
  if (raw) {
    p->ops.gofloat.is_cfa = 1;
    p->ops.basecurve.npoints = 1;
    p->ops.basecurve.points[0][0] = 0.50f;
    p->ops.basecurve.points[0][1] = 0.60f;
    for (int i = 0; i < 4; i++) p->ops.tolab.wb_coeffs[i] = 1.0f;
  } else {
    memcpy(p->ops.tolab.cam_to_xyz, SRGB_D65_43, sizeof(SRGB_D65_43));
    memcpy(p->ops.tolab.cam_to_xyz_normalized, SRGB_D65_43, sizeof(SRGB_D65_43));
    memcpy(p->ops.tolab.xyz_to_cam, XYZ_D65_34, sizeof(XYZ_D65_34));
    p->ops.tolab.wb_coeffs[0] = 1.0f; p->ops.tolab.wb_coeffs[1] = 1.0f;
    p->ops.tolab.wb_coeffs[2] = 1.0f; p->ops.tolab.wb_coeffs[3] = 0.0f;
  }
}

/* pipeline.rs:286-288 default_ops() for ImageSource::Other (used by the fast path only) */
static int ops_are_default_other(const orc_pipeline *p) {
  orc_pipeline d;
  orc_pipeline_defaults(&d, &p->image);
  const orc_ops *a = &p->ops, *b = &d.ops;
  if (memcmp(&a->gofloat, &b->gofloat, sizeof(a->gofloat))) return 0;
  if (strcmp(a->demosaic.cfa, b->demosaic.cfa)) return 0;
  if (a->rotatecrop.crop_top != 0 || a->rotatecrop.crop_right != 0 || a->rotatecrop.crop_bottom != 0 ||
      a->rotatecrop.crop_left != 0 || a->rotatecrop.rotation != 0) return 0;
  if (memcmp(&a->tolab, &b->tolab, sizeof(a->tolab))) return 0;
  if (a->basecurve.exposure != 0 || a->basecurve.npoints != 0) return 0;
  if (a->transform.rotation != 0 || a->transform.fliph || a->transform.flipv) return 0;
  return 1;
}

/* pipeline.rs:313-338 reset + forward walk + clamp + reverse walk */
void orc_pipeline_negotiate(orc_pipeline *p, size_t *fw, size_t *fh) {
  orc_rotatecrop_reset(&p->ops.rotatecrop);
  size_t width = p->image.width, height = p->image.height, w, h;
  size_t xywh[4];
  orc_gofloat_size_image(&p->ops.gofloat, width, height, xywh); /* gofloat.rs:64-67 */
  width = xywh[2]; height = xywh[3];
  /* demosaic: default transform_forward */
  orc_rotatecrop_transform_forward(&p->ops.rotatecrop, width, height, &w, &h); width = w; height = h;
  /* tolab, basecurve, fromlab, gamma: default */
  orc_transform_transform_forward(&p->ops.transform, width, height, &w, &h); width = w; height = h;
  orc_scaling_size(width, height, p->settings.maxwidth, p->settings.maxheight, &w, &h);
  width = w; height = h;
  if (fw) *fw = width;
  if (fh) *fh = height;
  orc_transform_transform_forward(&p->ops.transform, width, height, &w, &h); width = w; height = h; /* reverse == forward */
  orc_rotatecrop_transform_reverse(&p->ops.rotatecrop, width, height, &w, &h); width = w; height = h;
  /* demosaic, gofloat: no transform_reverse */
  p->settings.demosaic_width = width;
  p->settings.demosaic_height = height;
}

static __thread double g_timings[8];
void orc_pipeline_last_timings(double ms[8]) { memcpy(ms, g_timings, sizeof(g_timings)); }

#define STEP(idx, expr)                          \
  do {                                           \
    double t0 = now_ms();                        \
    orc_buffer *nb = (expr);                     \
    g_timings[idx] = now_ms() - t0;              \
    if (!nb) { if (buf) orc_buffer_free(buf); return NULL; } \
    if (nb != buf && buf) orc_buffer_free(buf);  \
    buf = nb;                                    \
  } while (0)

/* pipeline.rs:311-375 run (cache = None) */
orc_buffer *orc_pipeline_run(orc_pipeline *p) {
  orc_pipeline_negotiate(p, NULL, NULL);
  orc_buffer *buf = NULL;
  STEP(0, orc_gofloat_run(&p->ops.gofloat, &p->image));
  STEP(1, orc_demosaic_run(&p->ops.demosaic, &p->settings, buf));
  STEP(2, orc_rotatecrop_run(&p->ops.rotatecrop, buf));
  STEP(3, orc_tolab_run(&p->ops.tolab, buf));
  STEP(4, orc_basecurve_run(&p->ops.basecurve, buf));
  STEP(5, orc_fromlab_run(buf));
  STEP(6, orc_gamma_run(&p->settings, buf));
  STEP(7, orc_transform_run(&p->ops.transform, buf));
  return buf;
}

/* pipeline.rs:377-422 */
int orc_pipeline_output_8bit(orc_pipeline *p, uint8_t **data, size_t *w, size_t *h) {
  int other = p->image.kind == ORC_SRC_RGB8 || p->image.kind == ORC_SRC_RGB16;
  if (other && p->settings.use_fastpath && ops_are_default_other(p)) { /* :381-402 */
    size_t width = p->image.width, height = p->image.height, n = width * height * 3;
    uint8_t *rgb = (uint8_t *)malloc(n ? n : 1);
    if (p->image.kind == ORC_SRC_RGB8) memcpy(rgb, p->image.data, n);
    else { /* image 0.24 to_rgb8 of a 16-bit image: (v + 128) / 257 — external crate, unpinned */
      const uint16_t *s = (const uint16_t *)p->image.data;
      for (size_t i = 0; i < n; i++) rgb[i] = (uint8_t)(((uint32_t)s[i] + 128u) / 257u);
    }
    size_t nw, nh;
    orc_scaling_size(width, height, p->settings.maxwidth, p->settings.maxheight, &nw, &nh);
    if (nw != width || nh != height) {
      uint8_t *o = (uint8_t *)malloc(nw * nh * 3 + 1);
      orc_scale_down_srgb(rgb, width, height, nw, nh, o);
      free(rgb); rgb = o;
    }
    *data = rgb; *w = nw; *h = nh;
    return 0;
  }
  p->settings.linear = 0; /* :405 */
  orc_buffer *buf = orc_pipeline_run(p);
  if (!buf) return -1;
  size_t n = buf->width * buf->height * 3;
  uint8_t *img = (uint8_t *)malloc(n ? n : 1);
  for (size_t i = 0; i < n; i++) img[i] = orc_output8bit(buf->data[i]); /* :408-414 serial loop */
  *data = img; *w = buf->width; *h = buf->height;
  orc_buffer_free(buf);
  return 0;
}

/* pipeline.rs:424-469 */
int orc_pipeline_output_16bit(orc_pipeline *p, uint16_t **data, size_t *w, size_t *h) {
  int other = p->image.kind == ORC_SRC_RGB8 || p->image.kind == ORC_SRC_RGB16;
  if (other && p->settings.use_fastpath && ops_are_default_other(p)) { /* :428-449 */
    size_t width = p->image.width, height = p->image.height, n = width * height * 3;
    uint16_t *rgb = (uint16_t *)malloc((n ? n : 1) * 2);
    if (p->image.kind == ORC_SRC_RGB16) memcpy(rgb, p->image.data, n * 2);
    else { /* image 0.24 to_rgb16 of an 8-bit image: v * 257 — external crate, unpinned */
      const uint8_t *s = (const uint8_t *)p->image.data;
      for (size_t i = 0; i < n; i++) rgb[i] = (uint16_t)((uint16_t)s[i] * 257u);
    }
    size_t nw, nh;
    orc_scaling_size(width, height, p->settings.maxwidth, p->settings.maxheight, &nw, &nh);
    if (nw != width || nh != height) {
      uint16_t *o = (uint16_t *)malloc((nw * nh * 3 + 1) * 2);
      orc_scale_down_srgb16(rgb, width, height, nw, nh, o);
      free(rgb); rgb = o;
    }
    *data = rgb; *w = nw; *h = nh;
    return 0;
  }
  p->settings.linear = 1; /* :452 */
  orc_buffer *buf = orc_pipeline_run(p);
  if (!buf) return -1;
  size_t n = buf->width * buf->height * 3;
  uint16_t *img = (uint16_t *)malloc((n ? n : 1) * 2);
  for (size_t i = 0; i < n; i++) img[i] = orc_output16bit(buf->data[i]); /* :455-461 */
  *data = img; *w = buf->width; *h = buf->height;
  orc_buffer_free(buf);
  return 0;
}
