/*
 * oracle.h — CPU restatement of the pedrocr/imagepipe OpBuffer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (imagepipe_b200/libipb200.so) never links, includes or calls anything in oracle/.
 *
 * Every function restates one function of the reference (Rust, cannot be compiled
 * here: no rustc/cargo, path deps rawloader/multicache absent) and cites the
 * reference file:line it follows (paths relative to /root/reference).
 *
 * Parity pinning: the restatement is checked (tests/test_oracle_kats.py) against every
 * golden vector / known-answer test the reference's own test-suite holds for this path
 * (color_conversions.rs:337-611, curves.rs:164-189, transform.rs:167-278,
 * scaling.rs:188-203, rotatecrop.rs:180-312, tests/roundtrip_test.rs,
 * tests/maxsize_test.rs).  Functions the reference does not test itself
 * (gofloat::run_raw, demosaic::full, scaled_demosaic numerics, CFA::color_at from the
 * absent rawloader 0.37 crate) are "parity unpinned" by the reference and are pinned
 * only by tests/golden/: vectors from a second, independent restatement (scalar numpy-f32
 * loops written from the reference source, tests/golden/make_golden.py) that this oracle
 * reproduces bit for bit (tests/test_golden.py).
 *
 * Arithmetic: f32 everywhere, compiled -ffp-contract=off (Rust never contracts to FMA),
 * glibc cbrtf/powf/exp2f/sinf/cosf stand in for Rust std (which calls the same libm).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/buffer.rs:4-11 — interleaved row-major f32 */
typedef struct orc_buffer {
  size_t width, height, colors;
  int monochrome;
  float *data;
} orc_buffer;

/* src/ops/gofloat.rs:4-12 */
typedef struct orc_gofloat {
  size_t crop_top, crop_right, crop_bottom, crop_left;
  int is_cfa;
  float blacklevels[4];
  float whitelevels[4];
} orc_gofloat;

/* src/ops/demosaic.rs:4-6 (pattern string of rawloader::CFA, ≤144 chars) */
typedef struct orc_demosaic {
  char cfa[148];
} orc_demosaic;

/* src/ops/rotatecrop.rs:10-18 */
typedef struct orc_rotatecrop {
  float crop_top, crop_right, crop_bottom, crop_left, rotation;
  float input_ratio;
  int has_output_size;
  size_t output_width, output_height;
} orc_rotatecrop;

/* src/ops/colorspaces.rs:5-10 */
typedef struct orc_tolab {
  float cam_to_xyz[3][4];
  float cam_to_xyz_normalized[3][4];
  float xyz_to_cam[4][3];
  float wb_coeffs[4];
} orc_tolab;

/* src/ops/curves.rs:6-9 */
#define ORC_MAX_CURVE_POINTS 32
typedef struct orc_basecurve {
  float exposure;
  size_t npoints;
  float points[ORC_MAX_CURVE_POINTS][2];
} orc_basecurve;

/* src/ops/transform.rs:7-19 */
enum { ORC_ROT_NORMAL = 0, ORC_ROT_90 = 1, ORC_ROT_180 = 2, ORC_ROT_270 = 3 };
typedef struct orc_transform {
  int rotation;
  int fliph, flipv;
} orc_transform;

/* src/pipeline.rs:110-118 */
typedef struct orc_settings {
  size_t maxwidth, maxheight, demosaic_width, demosaic_height;
  int linear;
  int use_fastpath;
} orc_settings;

/* src/pipeline.rs:47-50 ImageSource; rawloader RawImage{width,height,cpp,data}. */
enum { ORC_SRC_RAW_U16 = 0, ORC_SRC_RAW_F32 = 1, ORC_SRC_RGB8 = 2, ORC_SRC_RGB16 = 3 };
typedef struct orc_source {
  int kind;
  size_t width, height, cpp;
  const void *data;
} orc_source;

/* src/pipeline.rs:154-164 */
typedef struct orc_ops {
  orc_gofloat gofloat;
  orc_demosaic demosaic;
  orc_rotatecrop rotatecrop;
  orc_tolab tolab;
  orc_basecurve basecurve;
  orc_transform transform;
} orc_ops;

typedef struct orc_pipeline {
  orc_source image;
  orc_settings settings;
  orc_ops ops;
} orc_pipeline;

/* rawloader 0.37 CFA (src not in tree; restated from its published behaviour) */
typedef struct orc_cfa {
  size_t width, height;
  uint8_t pattern[48][48];
} orc_cfa;

/* ---- buffer ---- */
orc_buffer *orc_buffer_new(size_t w, size_t h, size_t colors, int mono);
orc_buffer *orc_buffer_from(size_t w, size_t h, size_t colors, int mono, const float *data);
orc_buffer *orc_buffer_clone(const orc_buffer *b);
void orc_buffer_free(orc_buffer *b);
void orc_set_threads(int n); /* 0 = all */
int orc_get_threads(void);

/* ---- color_conversions.rs ---- */
void orc_matrices(float srgb_d65_33[9], float xyz_d65_33[9]);
const float *orc_lut_xyz_lab(void);       /* 8193 entries */
const float *orc_lut_srgb_reverse(void);  /* 8193 entries */
const float *orc_lut_srgb_transform(void);/* 8193 entries */
float orc_expand_srgb_gamma(float v);
float orc_apply_srgb_gamma(float v);
void orc_xyz_to_lab(float x, float y, float z, float out[3]);
void orc_lab_to_xyz(float l, float a, float b, float out[3]);
void orc_camera_to_lab(const float mul[4], const float cmatrix[12], const float pixin[4], float out[3]);
void orc_lab_to_rgb(const float rgbmatrix[9], const float pixin[3], float out[3]);
float orc_input8bit(uint8_t v);
float orc_input16bit(uint16_t v);
uint8_t orc_output8bit(float v);
uint16_t orc_output16bit(float v);

/* ---- rawloader::CFA ---- */
int orc_cfa_new(orc_cfa *cfa, const char *pattern);
size_t orc_cfa_color_at(const orc_cfa *cfa, size_t row, size_t col);

/* ---- scaling.rs ---- */
float orc_calculate_scale(size_t w, size_t h, size_t maxw, size_t maxh);
void orc_scaling_size(size_t w, size_t h, size_t maxw, size_t maxh, size_t *ow, size_t *oh);
void orc_transform_buffer_f32(const float *src, size_t width, size_t height,
                              const long topleft[2], const long topright[2], const long bottomleft[2],
                              size_t nwidth, size_t nheight, size_t components,
                              const orc_cfa *cfa, float *out);
void orc_transform_buffer_u8(const uint8_t *src, size_t width, size_t height,
                             const long topleft[2], const long topright[2], const long bottomleft[2],
                             size_t nwidth, size_t nheight, size_t components, uint8_t *out);
void orc_transform_buffer_u16(const uint16_t *src, size_t width, size_t height,
                              const long topleft[2], const long topright[2], const long bottomleft[2],
                              size_t nwidth, size_t nheight, size_t components, uint16_t *out);
orc_buffer *orc_scaled_demosaic(const orc_cfa *cfa, const orc_buffer *buf, size_t nw, size_t nh);
orc_buffer *orc_scale_down_opbuf(const orc_buffer *buf, size_t nw, size_t nh);
/* lanczos.c — EXTENSION with no reference implementation (scaling.rs:101-103 is a FIXME): parity unpinned */
orc_buffer *orc_lanczos_resize(const orc_buffer *buf, size_t nw, size_t nh, int a);
void orc_lanczos_weights(size_t n_in, size_t n_out, int a, int *start, int *count, float *w, size_t *ksize_out);
void orc_scale_down_srgb(const uint8_t *src, size_t w, size_t h, size_t nw, size_t nh, uint8_t *out);
void orc_scale_down_srgb16(const uint16_t *src, size_t w, size_t h, size_t nw, size_t nh, uint16_t *out);

/* ---- curves.rs SplineFunc ---- */
typedef struct orc_spline {
  size_t n;                       /* number of points */
  float x[ORC_MAX_CURVE_POINTS + 2], y[ORC_MAX_CURVE_POINTS + 2];
  float c1[ORC_MAX_CURVE_POINTS + 2], c2[ORC_MAX_CURVE_POINTS + 2], c3[ORC_MAX_CURVE_POINTS + 2];
  size_t nseg;                    /* c3s.len() */
} orc_spline;
void orc_spline_new(orc_spline *s, const float (*p)[2], size_t n);
float orc_spline_interpolate(const orc_spline *s, float val);

/* ---- ops: each mirrors ImageOp::run; return value is a NEW buffer unless it is the
 * input pointer itself (the reference returns the same Arc for pass-through). ---- */
void orc_gofloat_size_image(const orc_gofloat *op, size_t ow, size_t oh, size_t out_xywh[4]);
orc_buffer *orc_gofloat_run(const orc_gofloat *op, const orc_source *img);
orc_buffer *orc_demosaic_full(const orc_cfa *cfa, const orc_buffer *buf);
orc_buffer *orc_demosaic_run(const orc_demosaic *op, const orc_settings *s, orc_buffer *buf);
orc_buffer *orc_rotatecrop_run(const orc_rotatecrop *op, orc_buffer *buf);
void orc_rotatecrop_transform_forward(orc_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh);
void orc_rotatecrop_transform_reverse(orc_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh);
void orc_rotatecrop_reset(orc_rotatecrop *op);
orc_buffer *orc_tolab_run(const orc_tolab *op, orc_buffer *buf);
orc_buffer *orc_basecurve_run(const orc_basecurve *op, orc_buffer *buf);
orc_buffer *orc_fromlab_run(orc_buffer *buf);
orc_buffer *orc_gamma_run(const orc_settings *s, orc_buffer *buf);
void orc_orientation_flips(const orc_transform *op, int flips[3]);
orc_buffer *orc_rotate_buffer(const orc_buffer *buf, int transpose, int flip_x, int flip_y);
orc_buffer *orc_transform_run(const orc_transform *op, orc_buffer *buf);
void orc_transform_transform_forward(const orc_transform *op, size_t w, size_t h, size_t *ow, size_t *oh);

/* ---- pipeline.rs ---- */
void orc_pipeline_defaults(orc_pipeline *p, const orc_source *img); /* PipelineOps::new for literal metadata */
void orc_pipeline_negotiate(orc_pipeline *p, size_t *final_w, size_t *final_h);
orc_buffer *orc_pipeline_run(orc_pipeline *p);
/* per-stage timings of the last orc_pipeline_run on this thread, ms: gofloat..transform */
void orc_pipeline_last_timings(double ms[8]);
int orc_pipeline_output_8bit(orc_pipeline *p, uint8_t **data, size_t *w, size_t *h);
int orc_pipeline_output_16bit(orc_pipeline *p, uint16_t **data, size_t *w, size_t *h);
void orc_free(void *p);

/* ---- KATs restated from the reference's own tests; each returns #mismatches ---- */
long orc_kat_roundtrip_8bit(void);
long orc_kat_roundtrip_16bit(void);
long orc_kat_roundtrip_8bit_gamma(void);
long orc_kat_roundtrip_16bit_gamma(void);
long orc_kat_roundtrip_8bit_lab_xyz(void);
long orc_kat_roundtrip_8bit_lab_rgb(void);
long orc_kat_roundtrip_8bit_lab_rgb_gamma(void);
long orc_kat_roundtrip_16bit_lab_xyz(void);
long orc_kat_roundtrip_16bit_lab_rgb(void);
long orc_kat_roundtrip_16bit_lab_rgb_gamma(void);
long orc_kat_rotatecrop_roundtrip_transform(void);
long orc_kat_rotatecrop_roundtrip_transform_rotation(void);

#ifdef __cplusplus
}
#endif
#endif
