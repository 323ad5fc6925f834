/*
 * kats.c — the reference's own known-answer tests for the hot path, re-expressed against
 * the oracle restatement.  TEST INFRASTRUCTURE ONLY.  Each function returns the number of
 * mismatches (0 == the reference's assert would hold everywhere).
 * Citations are to /root/reference/src/color_conversions.rs and src/ops/rotatecrop.rs.
 */
#include "oracle.h"
#include <string.h>

static float roundtrip_gamma(float v) { return orc_apply_srgb_gamma(orc_expand_srgb_gamma(v)); } /* :385-388 */

/* :337-342 (0..u8::MAX excludes 255) */
long orc_kat_roundtrip_8bit(void) {
  long bad = 0;
  for (int i = 0; i < 255; i++) bad += (orc_output8bit(orc_input8bit((uint8_t)i)) != i);
  return bad;
}
/* :344-349 */
long orc_kat_roundtrip_16bit(void) {
  long bad = 0;
  for (int i = 0; i < 65535; i++) bad += (orc_output16bit(orc_input16bit((uint16_t)i)) != i);
  return bad;
}
/* :390-395 */
long orc_kat_roundtrip_8bit_gamma(void) {
  long bad = 0;
  for (int i = 0; i < 255; i++) bad += (orc_output8bit(roundtrip_gamma(orc_input8bit((uint8_t)i))) != i);
  return bad;
}
/* :397-402 */
long orc_kat_roundtrip_16bit_gamma(void) {
  long bad = 0;
  for (int i = 0; i < 65535; i++) bad += (orc_output16bit(roundtrip_gamma(orc_input16bit((uint16_t)i))) != i);
  return bad;
}

/* :420-440 */
long orc_kat_roundtrip_8bit_lab_xyz(void) {
  long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int x = 0; x < 255; x++)
    for (int y = 0; y < 255; y++)
      for (int z = 0; z < 255; z++) {
        float lab[3], o[3];
        orc_xyz_to_lab(orc_input8bit(x), orc_input8bit(y), orc_input8bit(z), lab);
        orc_lab_to_xyz(lab[0], lab[1], lab[2], o);
        bad += !(orc_output8bit(o[0]) == x && orc_output8bit(o[1]) == y && orc_output8bit(o[2]) == z);
      }
  return bad;
}

static void srgb_mats(float cm[12], float rgbm[9]) {
  float s[9];
  orc_matrices(s, rgbm);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) cm[i * 4 + j] = s[i * 3 + j];
    cm[i * 4 + 3] = 0.0f;
  }
}

/* :442-463 */
long orc_kat_roundtrip_8bit_lab_rgb(void) {
  long bad = 0;
  float cm[12], rgbm[9];
  srgb_mats(cm, rgbm);
  const float mul[4] = {1, 1, 1, 1};
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int r = 0; r < 255; r++)
    for (int g = 0; g < 255; g++)
      for (int b = 0; b < 255; b++) {
        float px[4] = {orc_input8bit(r), orc_input8bit(g), orc_input8bit(b), 0.0f}, lab[3], o[3];
        orc_camera_to_lab(mul, cm, px, lab);
        orc_lab_to_rgb(rgbm, lab, o);
        bad += !(orc_output8bit(o[0]) == r && orc_output8bit(o[1]) == g && orc_output8bit(o[2]) == b);
      }
  return bad;
}

/* :465-495 */
long orc_kat_roundtrip_8bit_lab_rgb_gamma(void) {
  long bad = 0;
  float cm[12], rgbm[9];
  srgb_mats(cm, rgbm);
  const float mul[4] = {1, 1, 1, 1};
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int r = 0; r < 255; r++)
    for (int g = 0; g < 255; g++)
      for (int b = 0; b < 255; b++) {
        float px[4] = {orc_expand_srgb_gamma(orc_input8bit(r)), orc_expand_srgb_gamma(orc_input8bit(g)),
                       orc_expand_srgb_gamma(orc_input8bit(b)), 0.0f};
        float lab[3], o[3];
        orc_camera_to_lab(mul, cm, px, lab);
        orc_lab_to_rgb(rgbm, lab, o);
        bad += !(orc_output8bit(orc_apply_srgb_gamma(o[0])) == r && orc_output8bit(orc_apply_srgb_gamma(o[1])) == g &&
                 orc_output8bit(orc_apply_srgb_gamma(o[2])) == b);
      }
  return bad;
}

/* :497-530 */
long orc_kat_roundtrip_16bit_lab_xyz(void) {
  long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int x = 0; x < 65535; x += 89)
    for (int y = 0; y < 65535; y += 97)
      for (int z = 0; z < 65535; z += 101) {
        float lab[3], o[3];
        orc_xyz_to_lab(orc_input16bit(x), orc_input16bit(y), orc_input16bit(z), lab);
        orc_lab_to_xyz(lab[0], lab[1], lab[2], o);
        bad += !(orc_output16bit(o[0]) == x && orc_output16bit(o[1]) == y && orc_output16bit(o[2]) == z);
        bad += !(orc_output8bit(o[0]) == (x >> 8) && orc_output8bit(o[1]) == (y >> 8) && orc_output8bit(o[2]) == (z >> 8));
      }
  return bad;
}

/* :532-565 */
long orc_kat_roundtrip_16bit_lab_rgb(void) {
  long bad = 0;
  float cm[12], rgbm[9];
  srgb_mats(cm, rgbm);
  const float mul[4] = {1, 1, 1, 1};
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int r = 0; r < 65535; r += 89)
    for (int g = 0; g < 65535; g += 97)
      for (int b = 0; b < 65535; b += 101) {
        float px[4] = {orc_input16bit(r), orc_input16bit(g), orc_input16bit(b), 0.0f}, lab[3], o[3];
        orc_camera_to_lab(mul, cm, px, lab);
        orc_lab_to_rgb(rgbm, lab, o);
        bad += !(orc_output16bit(o[0]) == r && orc_output16bit(o[1]) == g && orc_output16bit(o[2]) == b);
        bad += !(orc_output8bit(o[0]) == (r >> 8) && orc_output8bit(o[1]) == (g >> 8) && orc_output8bit(o[2]) == (b >> 8));
      }
  return bad;
}

/* :567-611 — 16-bit outputs allowed off by one (assert_offby ..., 1, 1), 8-bit exact */
long orc_kat_roundtrip_16bit_lab_rgb_gamma(void) {
  long bad = 0;
  float cm[12], rgbm[9];
  srgb_mats(cm, rgbm);
  const float mul[4] = {1, 1, 1, 1};
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int r = 0; r < 65535; r += 89)
    for (int g = 0; g < 65535; g += 97)
      for (int b = 0; b < 65535; b += 101) {
        float px[4] = {orc_expand_srgb_gamma(orc_input16bit(r)), orc_expand_srgb_gamma(orc_input16bit(g)),
                       orc_expand_srgb_gamma(orc_input16bit(b)), 0.0f};
        float lab[3], o[3];
        orc_camera_to_lab(mul, cm, px, lab);
        lab[0] = roundtrip_gamma(lab[0]);
        orc_lab_to_rgb(rgbm, lab, o);
        o[0] = orc_apply_srgb_gamma(o[0]); o[1] = orc_apply_srgb_gamma(o[1]); o[2] = orc_apply_srgb_gamma(o[2]);
        int in[3] = {r, g, b};
        for (int c = 0; c < 3; c++) {
          int v = orc_output16bit(o[c]);
          int lo = in[c] > 0 ? in[c] - 1 : 0, hi = in[c] < 65535 ? in[c] + 1 : 65535;
          bad += !(v >= lo && v <= hi);
          bad += !(orc_output8bit(o[c]) == (in[c] >> 8));
        }
      }
  return bad;
}

/* rotatecrop.rs:273-294 */
long orc_kat_rotatecrop_roundtrip_transform(void) {
  long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int dim = 0; dim < 10000; dim += 89) {
    orc_rotatecrop op;
    memset(&op, 0, sizeof(op));
    op.input_ratio = 1.0f;
    for (int c1 = 0; c1 < 65535; c1 += 97)
      for (int c2 = 0; c2 < 65535; c2 += 101) {
        op.crop_top = orc_input16bit(c1); op.crop_right = orc_input16bit(c1);
        op.crop_bottom = orc_input16bit(c2); op.crop_left = orc_input16bit(c2);
        size_t iw, ih, rw, rh;
        orc_rotatecrop_transform_reverse(&op, dim, dim, &iw, &ih);
        orc_rotatecrop_transform_forward(&op, iw, ih, &rw, &rh);
        bad += !(rw == (size_t)dim && rh == (size_t)dim);
      }
  }
  return bad;
}

/* rotatecrop.rs:296-312 */
long orc_kat_rotatecrop_roundtrip_transform_rotation(void) {
  long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
  for (int width = 0; width < 10000; width += 89) {
    orc_rotatecrop op;
    memset(&op, 0, sizeof(op));
    op.input_ratio = 1.0f;
    for (int height = 0; height < 10000; height += 97)
      for (int rot = 0; rot < 255; rot++) {
        op.rotation = orc_input8bit((uint8_t)rot);
        size_t a, b, c, d, e, f;
        orc_rotatecrop_transform_forward(&op, width, height, &a, &b);
        orc_rotatecrop_transform_reverse(&op, a, b, &c, &d);
        orc_rotatecrop_transform_forward(&op, c, d, &e, &f);
        bad += !(e == a && f == b);
      }
  }
  return bad;
}
