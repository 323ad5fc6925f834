// ipb_host.cu — host side of libipb200.so: the C ABI of include/ipb200.h.
//
// What lives here is the host logic the reference keeps around its pixel loops, restated for a device-resident
// OpBuffer: contexts and streams, ref-counted buffers (Arc<OpBuffer>), parameter preparation (white-balance
// normalisation, matrices, look-up tables, spline coefficients, CFA tables), the size negotiation of
// Pipeline::run, op-by-op execution, and the decision to run the fused raw->sRGB kernels.  No pixel is ever
// computed on the host: every entry point that produces pixels launches a CUDA kernel or fails.
// Host float arithmetic is compiled -ffp-contract=off; all of it is f32 like the reference's.
#include "../../include/ipb200.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "ipb_internal.h"
#include "ipb_scaled.cuh"
#include "ipb_spec.h"

using namespace ipb;

// ------------------------------------------------------------------------------------------------ objects

struct ipb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  float2 *lut_lab = nullptr, *lut_gamma = nullptr, *lut_rev = nullptr;  // device {v, dv} tables
  float2 *lut_gamma8 = nullptr;  // device {threshold, base} table: output8bit(apply_srgb_gamma(v)) per segment
  float *cbrt_tab = nullptr;     // device table of cbrtf(v) for every float v in (1.0, 1.5]
  // Lanczos tap tables already on the device, most recently used last (ipb_lanczos_resize)
  struct LzTab { size_t n_in, n_out; int a, ksize; int *start, *count; float *w; std::vector<int> hstart, hcount; };
  std::vector<LzTab> lz_tabs;
  std::string err;
  unsigned long long launches = 0;
  // Stream-ordered allocations come from a pool of the context that keeps its memory between frames (the device's
  // default pool hands everything back to the driver at every synchronisation: a 384 MB OpBuffer then costs
  // milliseconds to map again).  Like the reference's allocator, it holds on to what a pipeline run needed.
  cudaMemPool_t pool = nullptr;
  // speculative 8-bit kernel (ipb_spec.cu): self-test results, device tables, the parameter set they were built for
  bool rc_cached = false;           // ranges_rc_exact: result for the last (black, range) pair
  float rc_black = 0.0f, rc_range = 0.0f;
  int rc_exact = 0;
  int spec_ok = 0;                  // the shared window starts where the gamma table's addressing assumes it does
  float mufu_cbrt_err = 1.0f;       // measured max relative error of the XU-pipe cube root (every float in [2^-8, 4])
  int spec_threads = 512;           // CTA size of k_spec8 (512: two CTAs per SM, 1024: one)
  float spec_delta_override = 0.0f; // tests: force the bound (0 = the certified one)
  uint32_t *spec_g8a = nullptr;
  float2 *spec_stab = nullptr;
  float *spec_thr = nullptr;        // the 255 exact thresholds of the 8-bit gamma step function
  unsigned long long *spec_stats = nullptr;   // 8 counters, see SpecParams::stats
  void *spec_stage = nullptr;       // pinned staging for table uploads (capturable copies)
  SpecTables spec_tab{};
  int spec_tab_state = 0;           // 0: none, 1: valid for spec_key, -1: spec_key cannot take the speculative path
  std::vector<unsigned char> spec_key;
  // copy streams + events of the chunk-pipelined host<->device path (created on first use)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  std::vector<cudaEvent_t> events;
};

struct ipb_buffer {
  std::atomic<int> refcnt{1};
  ipb_ctx *ctx = nullptr;
  size_t width = 0, height = 0, colors = 0;
  int monochrome = 0;
  float *dptr = nullptr;
  bool owned = true;
};

struct ipb_pipeline {
  ipb_ctx *ctx = nullptr;
  ipb_source image{};
  ipb_ops ops{};
  ipb_settings settings{};
  int fused = 1;
  int use_tma = 1;
  unsigned long long source_gen = 0;  // bumped by set_source: part of the cache key (a refilled buffer is a new image)
  int spec = 1;      // 8-bit output of RGB Bayer frames through the speculative kernel (results identical; 0: k_fused_full)
  // ipb_pipeline_output_8bit_batch in flight: frames of the batch, source rows / output bytes from one to the next, and
  // whether the launch took them all (else the caller runs them one by one)
  size_t batch_n = 0, batch_src_rows = 0, batch_out_bytes = 0;
  bool batch_done = false;
  int band_mb = 16;  // host<->device paths: band size of the overlapped H2D / kernel / D2H schedule (0 = no bands)
  // golevel_rc_exact() result for the last (black, range) pair: the check walks all 65536 samples
  bool rc_cached = false;
  float rc_black = 0.0f, rc_range = 0.0f;
  int rc_exact = 0;
  // row-stripe source (multi-GPU / chunked transfers)
  bool has_stripe = false;
  ipb_stripe stripe{};
  ipb_source stripe_rows{};
  // device staging for host-resident sources and host destinations (grown on demand, reused across runs)
  void *stage_in = nullptr;
  size_t stage_in_bytes = 0;
  void *stage_out = nullptr;
  size_t stage_out_bytes = 0;
  // last run with a cache: first op executed (8 = everything came from the cache) and number of ops executed
  int last_startpos = 0, last_ops_run = 0;
};

// Pipeline::new_cache (pipeline.rs:257-260): size-bounded LRU of device OpBuffers keyed by the cumulative hash of
// the settings and the op parameters up to and including the op that produced the buffer (pipeline.rs:340-361).
struct ipb_cache {
  struct Key {
    uint64_t a = 0, b = 0;
    bool operator==(const Key &o) const { return a == o.a && b == o.b; }
  };
  struct Entry { Key key; ipb_buffer *buf; size_t bytes; };
  ipb_ctx *ctx = nullptr;
  size_t max_bytes = 0, bytes = 0;
  std::vector<Entry> lru;  // front = least recently used
  std::mutex mu;           // the reference shares one cache between pipelines on different threads
  unsigned long long hits = 0, misses = 0;
};

namespace {

thread_local std::string g_create_err;

int fail(ipb_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_err = buf;
  return code;
}

#define IPB_CUDA(ctx, call)                                                                        \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail((ctx), IPB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
#define IPB_LAUNCH(ctx, call)                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail((ctx), IPB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    (ctx)->launches++;                                                                             \
  } while (0)
// stream-ordered allocation from the context's pool
static inline cudaError_t ipb_malloc_async(ipb_ctx *ctx, void **p, size_t bytes) {
  return ctx->pool ? cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream) : cudaMallocAsync(p, bytes, ctx->stream);
}
#define IPB_TRY(call)            \
  do {                           \
    int rc_ = (call);            \
    if (rc_ != IPB_OK) return rc_; \
  } while (0)

// ------------------------------------------------------------------------------------------------ constants

// color_conversions.rs:1-39 — matrices, evaluated in f32 exactly as the reference evaluates them at start-up
struct HostTables {
  float srgb_d65_33[3][3];
  float xyz_d65_33[3][3];
  float srgb_d65_43[3][4];
  float xyz_d65_34[4][3];
  float lab[kLutEntries + 1], rev[kLutEntries + 1], fwd[kLutEntries + 1];  // 8193 entries each
};

float lab_f_host(float v) {  // color_conversions.rs:120-124
  float e = 216.0f / 24389.0f;
  float k = 24389.0f / 27.0f;
  if (v > e) return cbrtf(v);
  return (k * v + 16.0f) / 116.0f;
}
float srgb_rev_host(float v) {  // :126-132
  if (v < 0.04045f) return v / 12.92f;
  return powf((v + 0.055f) / 1.055f, 2.4f);
}
float srgb_fwd_host(float v) {  // :134-140
  if (v < 0.0031308f) return v * 12.92f;
  return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}

const HostTables &tables() {
  static HostTables T;
  static std::once_flag once;
  std::call_once(once, [] {
    const float s[3][3] = {{0.4124564f, 0.3575761f, 0.1804375f},
                           {0.2126729f, 0.7151522f, 0.0721750f},
                           {0.0193339f, 0.1191920f, 0.9503041f}};
    memcpy(T.srgb_d65_33, s, sizeof(s));
    const float(*m)[3] = T.srgb_d65_33;
    float(*o)[3] = T.xyz_d65_33;
    // color_conversions.rs:20-39 inverse()
    float invdet = 1.0f / (m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) -
                           m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]));
    o[0][0] = (m[1][1] * m[2][2] - m[2][1] * m[1][2]) * invdet;
    o[0][1] = -(m[0][1] * m[2][2] - m[0][2] * m[2][1]) * invdet;
    o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * invdet;
    o[1][0] = -(m[1][0] * m[2][2] - m[1][2] * m[2][0]) * invdet;
    o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * invdet;
    o[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * invdet;
    o[2][0] = (m[1][0] * m[2][1] - m[2][0] * m[1][1]) * invdet;
    o[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * invdet;
    o[2][2] = (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * invdet;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) { T.srgb_d65_43[i][j] = m[i][j]; T.xyz_d65_34[i][j] = o[i][j]; }
      T.srgb_d65_43[i][3] = 0.0f;
      T.xyz_d65_34[3][i] = 0.0f;
    }
    // TransformLookup::new(13, f) — color_conversions.rs:87-94
    const float maxv = (float)(kLutEntries - 1);
    for (int i = 0; i <= kLutEntries; i++) {
      float v = (float)i / maxv;
      T.lab[i] = lab_f_host(v);
      T.rev[i] = srgb_rev_host(v);
      T.fwd[i] = srgb_fwd_host(v);
    }
  });
  return T;
}

// rawloader::CFA::new (crate absent; call sites demosaic.rs:32-33,80,86, scaling.rs:110): pattern string of
// length 4 / 36 / 16 / 144 -> 2x2 / 6x6 / 2 wide x 8 high / 12x12, R G B E -> 0 1 2 3 (M -> 1, Y -> 3)
int parse_cfa(const char *pat, CfaDev *out) {
  size_t len = strnlen(pat, 147);
  int w, h;
  switch (len) {
    case 0: w = 0; h = 0; break;
    case 4: w = 2; h = 2; break;
    case 36: w = 6; h = 6; break;
    case 16: w = 2; h = 8; break;
    case 144: w = 12; h = 12; break;
    default: return IPB_ERR_BAD_CFA;
  }
  memset(out, 0, sizeof(*out));
  out->width = w;
  out->height = h;
  if (w == 0) return IPB_OK;
  uint8_t base[144];
  for (size_t i = 0; i < len; i++) {
    switch (pat[i]) {
      case 'R': base[i] = 0; break;
      case 'G': base[i] = 1; break;
      case 'B': base[i] = 2; break;
      case 'E': base[i] = 3; break;
      case 'M': base[i] = 1; break;
      case 'Y': base[i] = 3; break;
      default: return IPB_ERR_BAD_CFA;
    }
  }
  for (int r = 0; r < 48; r++)
    for (int c = 0; c < 48; c++) out->pat[r * 48 + c] = base[(r % h) * w + (c % w)];
  return IPB_OK;
}

// colorspaces.rs:12-27
void normalize_wbs(const float vals[4], float out[4]) {
  float unity = vals[1];
  for (int i = 0; i < 4; i++) out[i] = !std::isnormal(vals[i]) ? 1.0f : vals[i] / unity;
}

// SplineFunc::new — curves.rs:68-124.  Returns false where the reference would panic (fewer than two points).
bool build_spline(const ipb_basecurve *op, SplineDev *s) {
  memset(s, 0, sizeof(*s));
  if (op->npoints == 0 && fabsf(op->exposure) < 0.001f) return true;  // pass-through, n == 0 (curves.rs:34-36)
  const size_t n = op->npoints;
  float px[IPB_MAX_CURVE_POINTS], py[IPB_MAX_CURVE_POINTS];
  const float ex = exp2f(op->exposure);
  for (size_t i = 0; i < n; i++) { px[i] = op->points[i][0]; py[i] = op->points[i][1] * ex; }  // curves.rs:38-41
  int np = 0;
  if (n == 0 || (px[0] > 0.0f && py[0] > 0.0f)) { s->x[np] = 0.0f; s->y[np] = 0.0f; np++; }
  for (size_t i = 0; i < n; i++) { s->x[np] = px[i]; s->y[np] = py[i]; np++; }
  if (n == 0 || (px[n - 1] < 1.0f && py[n - 1] < 1.0f)) { s->x[np] = 1.0f; s->y[np] = 1.0f; np++; }
  if (np < 2) return false;
  float dxs[kMaxSplinePts], slopes[kMaxSplinePts];
  const int nd = np - 1;
  for (int i = 0; i < nd; i++) {
    float dx = s->x[i + 1] - s->x[i];
    float dy = s->y[i + 1] - s->y[i];
    dxs[i] = dx;
    slopes[i] = dy / dx;
  }
  int nc1 = 0;
  s->c1[nc1++] = slopes[0];
  for (int i = 0; i + 1 < nd; i++) {
    float m = slopes[i], next = slopes[i + 1];
    if (m * next <= 0.0f) {
      s->c1[nc1++] = 0.0f;
    } else {
      float dx = dxs[i], dxnext = dxs[i + 1];
      float common = dx + dxnext;
      s->c1[nc1++] = 3.0f * common / ((common + dxnext) / m + (common + dx) / next);
    }
  }
  s->c1[nc1++] = slopes[nd - 1];
  for (int i = 0; i + 1 < nc1; i++) {
    float c1 = s->c1[i], slope = slopes[i];
    float invdx = 1.0f / dxs[i];
    float common = c1 + s->c1[i + 1] - slope - slope;
    s->c2[i] = (slope - c1 - common) * invdx;
    s->c3[i] = common * invdx * invdx;
  }
  s->n = np;
  s->nseg = nc1 - 1;
  s->x_first = s->x[0]; s->y_first = s->y[0];
  s->x_last = s->x[np - 1]; s->y_last = s->y[np - 1];
  s->y_nan = s->y[(s->nseg - 1) / 2];
  return true;
}

// OpToLab::run parameter preparation — colorspaces.rs:89-101
void fill_tolab(ColorParams *P, const ipb_tolab *op, int monochrome) {
  const HostTables &T = tables();
  if (monochrome) {
    memcpy(P->cm, T.srgb_d65_43, sizeof(P->cm));
    P->mul[0] = P->mul[1] = P->mul[2] = P->mul[3] = 1.0f;
  } else {
    memcpy(P->cm, op->cam_to_xyz_normalized, sizeof(P->cm));
    normalize_wbs(op->wb_coeffs, P->mul);
  }
  memcpy(P->rgbm, T.xyz_d65_33, sizeof(P->rgbm));
  P->use_e = 1;
}

// scaling.rs:8-23
void scaling_total(size_t width, size_t height, size_t maxwidth, size_t maxheight, float *scale, size_t *ow,
                   size_t *oh) {
  if (maxwidth == 0 && maxheight == 0) { *scale = 1.0f; *ow = width; *oh = height; return; }
  float xscale = maxwidth == 0 ? 1.0f : (float)width / (float)maxwidth;
  float yscale = maxheight == 0 ? 1.0f : (float)height / (float)maxheight;
  auto f2u = [](float f) -> size_t { return f > 0.0f ? (size_t)f : 0; };
  if (yscale <= 1.0f && xscale <= 1.0f) { *scale = 1.0f; *ow = width; *oh = height; }
  else if (yscale > xscale) { *scale = yscale; *ow = f2u((float)width / yscale); *oh = maxheight; }
  else { *scale = xscale; *ow = maxwidth; *oh = f2u((float)height / xscale); }
}

// gofloat.rs:74-82
void size_image(const ipb_gofloat *op, size_t ow, size_t oh, size_t o[4]) {
  auto umin = [](size_t a, size_t b) { return a < b ? a : b; };
  o[0] = umin(op->crop_left, ow - 10);
  o[1] = umin(op->crop_top, oh - 10);
  o[2] = ow - umin(op->crop_left + op->crop_right, ow - 10);
  o[3] = oh - umin(op->crop_top + op->crop_bottom, oh - 10);
}

// ---- rotatecrop.rs:89-163
const float kRcEps = 1.0f / 1000000.0f;
const float kFracPi2 = 1.57079632679489661923132169163975144f;
bool rc_noop(const ipb_rotatecrop *op) {
  return fabsf(op->rotation) < kRcEps && fabsf(op->crop_top) < kRcEps && fabsf(op->crop_right) < kRcEps &&
         fabsf(op->crop_bottom) < kRcEps && fabsf(op->crop_left) < kRcEps;
}
long f2isize(float f) {
  if (f != f) return 0;
  if (f >= 9223372036854775808.0f) return 0x7fffffffffffffffL;
  if (f <= -9223372036854775808.0f) return (long)0x8000000000000000UL;
  return (long)f;
}
size_t f2usize(float f) {
  if (!(f > 0.0f)) return 0;
  if (f >= 18446744073709551616.0f) return (size_t)-1;
  return (size_t)f;
}
void rc_rotate_point_reverse(const ipb_rotatecrop *op, float x, float y, float width, float height, float swidth,
                             float sheight, long out[2]) {
  if (op->rotation < kRcEps) { out[0] = f2isize(x); out[1] = f2isize(y); return; }
  float angle = kFracPi2 * (op->rotation > 1.0f ? 1.0f : op->rotation);
  float sn = sinf(angle), cs = cosf(angle);
  float tx = x - (width / 2.0f), ty = y - (height / 2.0f);
  float nx = tx * cs + ty * sn + (swidth / 2.0f);
  float ny = -tx * sn + ty * cs + (sheight / 2.0f);
  out[0] = f2isize(nx);
  out[1] = f2isize(ny);
}
void rc_calc_size(const ipb_rotatecrop *op, size_t owidth, size_t oheight, bool reverse, size_t *ow, size_t *oh) {
  if (rc_noop(op)) { *ow = owidth; *oh = oheight; return; }
  float width = (float)owidth, height = (float)oheight;
  if (!(reverse || op->rotation < kRcEps)) {
    float angle = kFracPi2 * (op->rotation > 1.0f ? 1.0f : op->rotation);
    float sn = sinf(angle), cs = cosf(angle);
    float w2 = width * cs + height * sn, h2 = width * sn + height * cs;
    width = w2;
    height = h2;
  }
  float nwidth, nheight;
  {
    float ratio = 1.0f - op->crop_left - op->crop_right;
    nwidth = reverse ? roundf(width / ratio) : roundf(width * ratio);
    if (ratio < kRcEps || nwidth < 1.0f) { *ow = owidth; *oh = oheight; return; }
  }
  {
    float ratio = 1.0f - op->crop_top - op->crop_bottom;
    nheight = reverse ? roundf(height / ratio) : roundf(height * ratio);
    if (ratio < kRcEps || nheight < 1.0f) { *ow = owidth; *oh = oheight; return; }
  }
  if (!(!reverse || op->rotation < kRcEps)) {
    float angle = kFracPi2 * (op->rotation > 1.0f ? 1.0f : op->rotation);
    float sn = sinf(angle), cs = cosf(angle);
    float w2 = roundf(nheight / (sn + (cs / op->input_ratio)));
    float h2 = roundf(w2 / op->input_ratio);
    nwidth = w2;
    nheight = h2;
  }
  *ow = f2usize(nwidth);
  *oh = f2usize(nheight);
}

// rawloader Orientation::to_flips for the four base rotations, XOR the user flips (transform.rs:56-66);
// the table is pinned by the eight golden bitmaps of transform.rs:167-278
void orientation_flips(const ipb_transform *op, int f[3]) {
  switch (op->rotation) {
    case IPB_ROT_90: f[0] = 1; f[1] = 0; f[2] = 1; break;
    case IPB_ROT_180: f[0] = 0; f[1] = 1; f[2] = 1; break;
    case IPB_ROT_270: f[0] = 1; f[1] = 1; f[2] = 0; break;
    default: f[0] = 0; f[1] = 0; f[2] = 0; break;
  }
  f[1] ^= (op->fliph != 0);
  f[2] ^= (op->flipv != 0);
}

// Is the 3-instruction reciprocal division of the device bit-identical to (v - black) / range for every
// possible u16 sample?  (gofloat.rs:127)
bool golevel_rc_exact(float black, float range, float rc) {
  if (!std::isfinite(rc) || !std::isnormal(range)) return false;
  for (int v = 0; v < 65536; v++) {
    float num = (float)v - black;
    float q = num * rc;
    float r = fmaf(-q, range, num);
    float q2 = fmaf(r, rc, q);
    float ref = num / range;
    if (memcmp(&q2, &ref, 4) != 0) return false;
  }
  return true;
}


// ---- 8-bit gamma threshold table.
// g(v) = output8bit(SRGB_GAMMA_TRANSFORM.lookup(v)) for v in [0,1] (gamma.rs:21 + pipeline.rs:408-414) is a
// non-decreasing step function of v: inside table segment `key` every operation of the lerp
// (pos = v*8191, a = pos - key, t[key] + a*dt with dt >= 0) and of output8bit (v*256, clamp, trunc) is monotone
// under round-to-nearest, and the segment ends meet because t[key] + dt == t[key+1] exactly (Sterbenz).  One
// segment spans at most 256*12.92/8191 < 1 output steps, so g is known from {g(first v of the segment), the
// smallest v where it steps up}.  The device evaluates base + (v >= threshold): identical bytes, ~half the
// instructions.  Everything below is verified while building; on any violation the table is not used.
float gamma_lerp_host(const float *t, float v) {
  float pos = v * (float)(kLutEntries - 1);
  float base = truncf(pos);
  int key = (int)base;
  float a = pos - base;
  return t[key] + a * (t[key + 1] - t[key]);
}
uint32_t out8_host(float v) {  // color_conversions.rs:323-325
  float t = v * 256.0f;
  t = t > 0.0f ? t : 0.0f;  // max(0): NaN -> 0
  t = t < 255.0f ? t : 255.0f;
  return (uint32_t)t;
}
uint32_t g8_host(const float *t, float v) { return out8_host(gamma_lerp_host(t, v)); }
float f_from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
uint32_t bits_of(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
int key_of(float v) { return (int)truncf(v * (float)(kLutEntries - 1)); }

struct Gamma8Entry { float thr; uint32_t base; };

bool build_gamma8(const float *t, std::vector<Gamma8Entry> *out) {
  out->assign(kLutEntries, Gamma8Entry{2.0f, 0u});
  const uint32_t one = bits_of(1.0f);
  // first float (by bit pattern; non-negative floats order like their bits) of every segment
  std::vector<uint32_t> first(kLutEntries + 1);
  for (int key = 0; key < kLutEntries; key++) {
    uint32_t lo = 0, hi = one;  // smallest bits b with key_of(b) >= key; key_of is monotone in b
    while (lo < hi) {
      uint32_t mid = lo + (hi - lo) / 2;
      if (key_of(f_from_bits(mid)) >= key) hi = mid; else lo = mid + 1;
    }
    first[key] = lo;
  }
  first[kLutEntries] = one + 1;
  uint32_t prev_top = 0;
  for (int key = 0; key < kLutEntries; key++) {
    const uint32_t b0 = first[key], b1 = first[key + 1] - 1;  // segment = bit patterns [b0, b1]
    if (b1 < b0 || key_of(f_from_bits(b0)) != key || key_of(f_from_bits(b1)) != key) return false;
    const uint32_t base = g8_host(t, f_from_bits(b0)), top = g8_host(t, f_from_bits(b1));
    if (base < prev_top || top < base || top > base + 1) return false;
    prev_top = top;
    Gamma8Entry e{2.0f, base};
    if (top == base + 1) {
      uint32_t lo = b0 + 1, hi = b1;  // smallest b with g8 == top
      while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (g8_host(t, f_from_bits(mid)) > base) hi = mid; else lo = mid + 1;
      }
      e.thr = f_from_bits(lo);
    }
    (*out)[key] = e;
  }
  return true;
}

std::vector<Gamma8Entry> g_gamma8;  // built once per process by the first context
bool g_gamma8_ok = false;

int upload_lut(ipb_ctx *ctx, const float *t, float2 **out) {
  std::vector<float2> host(kLutEntries);
  for (int i = 0; i < kLutEntries; i++) host[i] = make_float2(t[i], t[i + 1] - t[i]);
  IPB_CUDA(ctx, cudaMalloc((void **)out, kLutEntries * sizeof(float2)));
  IPB_CUDA(ctx, cudaMemcpy(*out, host.data(), kLutEntries * sizeof(float2), cudaMemcpyHostToDevice));
  return IPB_OK;
}

int new_buffer(ipb_ctx *ctx, size_t w, size_t h, size_t colors, int mono, bool zero, ipb_buffer **out) {
  ipb_buffer *b = new (std::nothrow) ipb_buffer();
  if (!b) return fail(ctx, IPB_ERR_NOMEM, "out of host memory");
  b->ctx = ctx; b->width = w; b->height = h; b->colors = colors; b->monochrome = mono;
  size_t bytes = w * h * colors * sizeof(float);
  cudaError_t e = ipb_malloc_async(ctx, (void **)&b->dptr, bytes ? bytes : 4);
  if (e == cudaSuccess && zero && bytes) e = cudaMemsetAsync(b->dptr, 0, bytes, ctx->stream);
  if (e != cudaSuccess) {
    delete b;
    return fail(ctx, e == cudaErrorMemoryAllocation ? IPB_ERR_NOMEM : IPB_ERR_CUDA, "buffer alloc %zux%zux%zu: %s", w, h,
                colors, cudaGetErrorString(e));
  }
  *out = b;
  return IPB_OK;
}

int enter(ipb_ctx *ctx) {
  if (!ctx) return fail(nullptr, IPB_ERR_INVALID, "null context");
  IPB_CUDA(ctx, cudaSetDevice(ctx->device));
  return IPB_OK;
}

size_t src_elem_size(int kind) {
  switch (kind) {
    case IPB_SRC_RAW_U16: return 2;
    case IPB_SRC_RAW_F32: return 4;
    case IPB_SRC_RGB8: return 1;
    default: return 2;
  }
}
size_t src_cpp(const ipb_source *s) { return (s->kind == IPB_SRC_RGB8 || s->kind == IPB_SRC_RGB16) ? 3 : s->cpp; }

// device view of a source: the pointer itself when on_device, otherwise a stream-ordered temporary copy
struct DevSrc {
  const void *ptr = nullptr;
  void *tmp = nullptr;
};
int device_source(ipb_ctx *ctx, const ipb_source *img, DevSrc *d) {
  if (!img->data) return fail(ctx, IPB_ERR_INVALID, "source has no data");
  if (img->on_device) { d->ptr = img->data; return IPB_OK; }
  size_t bytes = img->width * img->height * src_cpp(img) * src_elem_size(img->kind);
  IPB_CUDA(ctx, ipb_malloc_async(ctx, &d->tmp, bytes ? bytes : 4));
  IPB_CUDA(ctx, cudaMemcpyAsync(d->tmp, img->data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  d->ptr = d->tmp;
  return IPB_OK;
}
void release_source(ipb_ctx *ctx, DevSrc *d) {
  if (d->tmp) cudaFreeAsync(d->tmp, ctx->stream);
  d->tmp = nullptr;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ context

extern "C" {

int ipb_version(void) { return IPB_VERSION; }

static int ensure_cbrt_table(ipb_ctx *ctx);

int ipb_ctx_create(int device, void *stream, ipb_ctx **out) {
  if (!out) return fail(nullptr, IPB_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, IPB_ERR_CUDA, "no usable CUDA device (%s); imagepipe-b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev) return fail(nullptr, IPB_ERR_INVALID, "device %d out of range (%d present)", device, ndev);
  ipb_ctx *ctx = new (std::nothrow) ipb_ctx();
  if (!ctx) return fail(nullptr, IPB_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  auto bail = [&](int rc) { g_create_err = ctx->err; ipb_ctx_destroy(ctx); return rc; };
  if ((e = cudaSetDevice(device)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(IPB_ERR_CUDA); }
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      return bail(IPB_ERR_CUDA);
    }
    ctx->own_stream = true;
  }
  {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      ctx->pool = nullptr;  // fall back to the device's default pool
      cudaGetLastError();
    }
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (ctx->sm_count <= 0) ctx->sm_count = 148;
  // keep freed blocks in the stream-ordered pool instead of returning them to the driver at every sync
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  const HostTables &T = tables();
  int rc;
  if ((rc = upload_lut(ctx, T.lab, &ctx->lut_lab)) != IPB_OK) return bail(rc);
  if ((rc = upload_lut(ctx, T.fwd, &ctx->lut_gamma)) != IPB_OK) return bail(rc);
  if ((rc = upload_lut(ctx, T.rev, &ctx->lut_rev)) != IPB_OK) return bail(rc);
  {
    static std::vector<Gamma8Entry> &g8 = g_gamma8;
    static bool &g8_ok = g_gamma8_ok;
    static std::once_flag once;
    std::call_once(once, [&] { if (g8.empty()) g8_ok = build_gamma8(T.fwd, &g8); });
    if (g8_ok) {
      static_assert(sizeof(Gamma8Entry) == sizeof(float2), "table entry layout");
      if ((e = cudaMalloc((void **)&ctx->lut_gamma8, kLutEntries * sizeof(float2))) != cudaSuccess ||
          (e = cudaMemcpy(ctx->lut_gamma8, g8.data(), kLutEntries * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) {
        ctx->err = cudaGetErrorString(e);
        return bail(IPB_ERR_CUDA);
      }
    }
  }
  // built here rather than at the first fused launch, so that a first launch may sit inside a stream capture
  if (ensure_cbrt_table(ctx) != IPB_OK) {
    g_create_err = ctx->err;
    ipb_ctx_destroy(ctx);
    return IPB_ERR_CUDA;
  }
  // speculative kernel: buffers for its tables, and its self-test (shared window base, XU-pipe cube root accuracy)
  {
    unsigned int *d_st = nullptr, h_st[2] = {0, 0};
    if ((e = cudaMalloc((void **)&ctx->spec_g8a, kSpecG8Entries * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc((void **)&ctx->spec_stab, kSpecSTabEntries * sizeof(float2))) != cudaSuccess ||
        (e = cudaMalloc((void **)&ctx->spec_thr, 256 * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc((void **)&ctx->spec_stats, 8 * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemset(ctx->spec_stats, 0, 8 * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaHostAlloc(&ctx->spec_stage, kSpecG8Entries * sizeof(uint32_t) + kSpecSTabEntries * sizeof(float2), cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaMalloc((void **)&d_st, sizeof(h_st))) != cudaSuccess ||
        (e = cudaMemset(d_st, 0, sizeof(h_st))) != cudaSuccess ||
        (e = launch_spec_selftest(ctx->stream, d_st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(h_st, d_st, sizeof(h_st), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
      ctx->err = std::string("speculative kernel self-test: ") + cudaGetErrorString(e);
      if (d_st) cudaFree(d_st);
      return bail(IPB_ERR_CUDA);
    }
    cudaFree(d_st);
    memcpy(&ctx->mufu_cbrt_err, &h_st[1], 4);
    if (g_gamma8_ok) {
      std::vector<float> thr;
      for (const Gamma8Entry &g : g_gamma8)
        if (g.thr <= 1.0f) thr.push_back(g.thr);
      if (thr.size() != 255) g_gamma8_ok = false;
      thr.resize(256, 2.0f);
      if ((e = cudaMemcpy(ctx->spec_thr, thr.data(), 256 * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
        ctx->err = cudaGetErrorString(e);
        return bail(IPB_ERR_CUDA);
      }
    }
    ctx->spec_ok = (h_st[0] == kSpecSmemBase && g_gamma8_ok && ctx->mufu_cbrt_err < 4.0e-6f) ? 1 : 0;
  }
  *out = ctx;
  return IPB_OK;
}

void ipb_ctx_destroy(ipb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->lut_lab) cudaFree(ctx->lut_lab);
  if (ctx->lut_gamma) cudaFree(ctx->lut_gamma);
  if (ctx->lut_rev) cudaFree(ctx->lut_rev);
  if (ctx->lut_gamma8) cudaFree(ctx->lut_gamma8);
  if (ctx->cbrt_tab) cudaFree(ctx->cbrt_tab);
  if (ctx->spec_g8a) cudaFree(ctx->spec_g8a);
  if (ctx->spec_stab) cudaFree(ctx->spec_stab);
  if (ctx->spec_thr) cudaFree(ctx->spec_thr);
  if (ctx->spec_stats) cudaFree(ctx->spec_stats);
  if (ctx->spec_stage) cudaFreeHost(ctx->spec_stage);
  for (auto &t : ctx->lz_tabs) cudaFree(t.start);
  if (ctx->copy_in) { cudaStreamSynchronize(ctx->copy_in); cudaStreamDestroy(ctx->copy_in); }
  if (ctx->copy_out) { cudaStreamSynchronize(ctx->copy_out); cudaStreamDestroy(ctx->copy_out); }
  for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int ipb_ctx_set_stream(ipb_ctx *ctx, void *stream) {
  IPB_TRY(enter(ctx));
  if (ctx->own_stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
  }
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    IPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return IPB_OK;
}

int ipb_ctx_synchronize(ipb_ctx *ctx) {
  IPB_TRY(enter(ctx));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IPB_OK;
}

const char *ipb_last_error(const ipb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
unsigned long long ipb_ctx_launch_count(const ipb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int ipb_host_alloc(size_t bytes, void **out) {
  if (!out) return IPB_ERR_INVALID;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(nullptr, IPB_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
  return IPB_OK;
}
void ipb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int ipb_device_alloc(ipb_ctx *ctx, size_t bytes, void **out) {
  IPB_TRY(enter(ctx));
  if (!out) return fail(ctx, IPB_ERR_INVALID, "null out pointer");
  cudaError_t e = ipb_malloc_async(ctx, out, bytes ? bytes : 4);
  if (e != cudaSuccess)
    return fail(ctx, e == cudaErrorMemoryAllocation ? IPB_ERR_NOMEM : IPB_ERR_CUDA, "device alloc %zu: %s", bytes, cudaGetErrorString(e));
  return IPB_OK;
}
int ipb_device_free(ipb_ctx *ctx, void *dptr) {
  IPB_TRY(enter(ctx));
  if (dptr) IPB_CUDA(ctx, cudaFreeAsync(dptr, ctx->stream));
  return IPB_OK;
}
int ipb_device_upload(ipb_ctx *ctx, void *dptr, const void *host, size_t bytes) {
  IPB_TRY(enter(ctx));
  if (bytes && (!dptr || !host)) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (bytes) IPB_CUDA(ctx, cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IPB_OK;
}
int ipb_device_download(ipb_ctx *ctx, void *host, const void *dptr, size_t bytes) {
  IPB_TRY(enter(ctx));
  if (bytes && (!dptr || !host)) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (bytes) IPB_CUDA(ctx, cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IPB_OK;
}

// ------------------------------------------------------------------------------------------------ OpBuffer

int ipb_buffer_new(ipb_ctx *ctx, size_t width, size_t height, size_t colors, int monochrome, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!out) return fail(ctx, IPB_ERR_INVALID, "null out pointer");
  return new_buffer(ctx, width, height, colors, monochrome, true, out);
}

int ipb_buffer_upload(ipb_ctx *ctx, size_t width, size_t height, size_t colors, int monochrome, const float *host,
                      ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!out || (!host && width * height * colors)) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, width, height, colors, monochrome, false, &b));
  size_t bytes = width * height * colors * sizeof(float);
  if (bytes) {
    cudaError_t e = cudaMemcpyAsync(b->dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // the host array may be pageable and short-lived
    if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "upload: %s", cudaGetErrorString(e)); }
  }
  *out = b;
  return IPB_OK;
}

int ipb_buffer_wrap(ipb_ctx *ctx, size_t width, size_t height, size_t colors, int monochrome, void *dptr,
                    ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!out || !dptr) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  ipb_buffer *b = new (std::nothrow) ipb_buffer();
  if (!b) return fail(ctx, IPB_ERR_NOMEM, "out of host memory");
  b->ctx = ctx; b->width = width; b->height = height; b->colors = colors; b->monochrome = monochrome;
  b->dptr = (float *)dptr;
  b->owned = false;
  *out = b;
  return IPB_OK;
}

int ipb_buffer_download(ipb_ctx *ctx, const ipb_buffer *buf, float *host) {
  IPB_TRY(enter(ctx));
  if (!buf || !host) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  size_t bytes = buf->width * buf->height * buf->colors * sizeof(float);
  if (bytes) IPB_CUDA(ctx, cudaMemcpyAsync(host, buf->dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IPB_OK;
}

void ipb_buffer_retain(ipb_buffer *buf) {
  if (buf) buf->refcnt.fetch_add(1, std::memory_order_relaxed);
}
void ipb_buffer_release(ipb_buffer *buf) {
  if (!buf) return;
  if (buf->refcnt.fetch_sub(1, std::memory_order_acq_rel) == 1) {
    if (buf->owned && buf->dptr) {
      cudaSetDevice(buf->ctx->device);
      cudaFreeAsync(buf->dptr, buf->ctx->stream);
    }
    delete buf;
  }
}
size_t ipb_buffer_width(const ipb_buffer *buf) { return buf ? buf->width : 0; }
size_t ipb_buffer_height(const ipb_buffer *buf) { return buf ? buf->height : 0; }
size_t ipb_buffer_colors(const ipb_buffer *buf) { return buf ? buf->colors : 0; }
int ipb_buffer_monochrome(const ipb_buffer *buf) { return buf ? buf->monochrome : 0; }
void *ipb_buffer_device_ptr(const ipb_buffer *buf) { return buf ? buf->dptr : nullptr; }

// ------------------------------------------------------------------------------------------------ ImageOp::run

// golevel_rc_exact for the context's last (black, range) pair (the check walks all 65536 samples)
static bool ranges_rc_exact(ipb_ctx *ctx, float black, float range) {
  if (!ctx->rc_cached || memcmp(&ctx->rc_black, &black, 4) != 0 || memcmp(&ctx->rc_range, &range, 4) != 0) {
    ctx->rc_exact = golevel_rc_exact(black, range, 1.0f / range) ? 1 : 0;
    ctx->rc_black = black;
    ctx->rc_range = range;
    ctx->rc_cached = true;
  }
  return ctx->rc_exact != 0;
}

int ipb_gofloat_run(ipb_ctx *ctx, const ipb_gofloat *op, const ipb_source *image, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!op || !image || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (image->width < 10 || image->height < 10)
    return fail(ctx, IPB_ERR_INVALID, "gofloat: source smaller than 10x10 (size_image underflows, gofloat.rs:77-80)");
  size_t xywh[4];
  size_image(op, image->width, image->height, xywh);
  const size_t x = xywh[0], y = xywh[1], width = xywh[2], height = xywh[3];
  DevSrc src;
  IPB_TRY(device_source(ctx, image, &src));
  ipb_buffer *b = nullptr;
  int rc = IPB_OK;
  if (image->kind == IPB_SRC_RAW_U16 || image->kind == IPB_SRC_RAW_F32) {
    float mins[4], ranges[4];
    for (int i = 0; i < 4; i++) { mins[i] = op->blacklevels[i]; ranges[i] = op->whitelevels[i] - mins[i]; }  // gofloat.rs:86-89
    int mode;
    size_t colors;
    int mono = 0;
    if (image->cpp == 1 && !op->is_cfa) { mode = 0; colors = 4; mono = 1; }
    else if (image->cpp == 3) { mode = 1; colors = 4; }
    else { mode = 2; colors = image->cpp; }
    // the CFA / plain branch leaves elements at 0.0 only if the raster is shorter than its header says (gofloat.rs:126)
    const bool zero = mode == 2 && !gofloat_rows_cover(image->width * image->height * image->cpp, image->width, x, y, width,
                                                       height, image->cpp);
    rc = new_buffer(ctx, width, height, colors, mono, zero, &b);
    if (rc == IPB_OK) {
      cudaError_t e = launch_gofloat_raw(ctx->stream, image->kind == IPB_SRC_RAW_F32, src.ptr,
                                         image->width * image->height * image->cpp, image->width, x, y, width, height,
                                         image->cpp, mode, mins, ranges,
                                         mode == 2 && ranges_rc_exact(ctx, mins[0], ranges[0]) ? 1 : 0, b->dptr);
      if (e != cudaSuccess) rc = fail(ctx, IPB_ERR_CUDA, "gofloat kernel: %s", cudaGetErrorString(e));
      else ctx->launches++;
    }
  } else if (image->kind == IPB_SRC_RGB8 || image->kind == IPB_SRC_RGB16) {
    rc = new_buffer(ctx, width, height, 4, 0, false, &b);
    if (rc == IPB_OK) {
      cudaError_t e = launch_gofloat_other(ctx->stream, image->kind == IPB_SRC_RGB16, src.ptr, image->width, x, y, width,
                                           height, ctx->lut_rev, b->dptr);
      if (e != cudaSuccess) rc = fail(ctx, IPB_ERR_CUDA, "gofloat kernel: %s", cudaGetErrorString(e));
      else ctx->launches++;
    }
  } else {
    rc = fail(ctx, IPB_ERR_INVALID, "unknown source kind %d", image->kind);
  }
  release_source(ctx, &src);
  if (rc != IPB_OK) { if (b) ipb_buffer_release(b); return rc; }
  *out = b;
  return IPB_OK;
}

static int scale_opbuf(ipb_ctx *ctx, const CfaDev *cfa, ipb_buffer *in, size_t nw, size_t nh, size_t out_colors,
                       ipb_buffer **out) {
  if (nw == 0 || nh == 0) return fail(ctx, IPB_ERR_INVALID, "scale to an empty image");
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, nw, nh, out_colors, in->monochrome, false, &b));
  XformGeom g;
  g.tl[0] = 0; g.tl[1] = 0;
  g.tr[0] = (long)in->width - 1; g.tr[1] = 0;
  g.bl[0] = 0; g.bl[1] = (long)in->height - 1;
  g.width = in->width; g.height = in->height; g.nwidth = nw; g.nheight = nh; g.components = out_colors;
  cudaError_t e = launch_transform_f32(ctx->stream, g, cfa, in->dptr, b->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "scale kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  *out = b;
  return IPB_OK;
}

static float cfa_minscale(const CfaDev &cfa) {  // demosaic.rs:33-39
  switch (cfa.width) {
    case 2: return 2.0f;
    case 6: return 3.0f;
    case 8: return 2.0f;
    case 12: return 12.0f;
    default: return 2.0f;
  }
}

int ipb_demosaic_run(ipb_ctx *ctx, const ipb_demosaic *op, const ipb_settings *settings, ipb_buffer *in,
                     ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!op || !settings || !in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  const size_t nwidth = settings->demosaic_width, nheight = settings->demosaic_height;
  float scale;
  size_t sw, sh;
  scaling_total(in->width, in->height, nwidth, nheight, &scale, &sw, &sh);
  CfaDev cfa;
  if (parse_cfa(op->cfa, &cfa) != IPB_OK) return fail(ctx, IPB_ERR_BAD_CFA, "demosaic: bad CFA pattern \"%s\"", op->cfa);
  const float minscale = cfa_minscale(cfa);
  if (scale <= 1.0f && in->colors == 4) {  // demosaic.rs:41-43
    ipb_buffer_retain(in);
    *out = in;
    return IPB_OK;
  }
  if (in->colors == 4) return scale_opbuf(ctx, nullptr, in, nwidth, nheight, 4, out);  // :44-46
  if (cfa.width == 0) return fail(ctx, IPB_ERR_BAD_CFA, "demosaic: %zu-channel buffer without a CFA pattern", in->colors);
  if (in->colors != 1) return fail(ctx, IPB_ERR_BAD_COLORS, "demosaic: expected 1 channel, got %zu", in->colors);
  if (scale >= minscale) return scale_opbuf(ctx, &cfa, in, nwidth, nheight, 4, out);  // :47-50 scaled_demosaic
  ipb_buffer *full;                                                                   // :51-60
  IPB_TRY(new_buffer(ctx, in->width, in->height, 4, in->monochrome, false, &full));
  cudaError_t e = launch_demosaic_full(ctx->stream, cfa, in->dptr, in->width, in->height, full->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(full); return fail(ctx, IPB_ERR_CUDA, "demosaic kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  if (scale > 1.0f) {
    int rc = scale_opbuf(ctx, nullptr, full, nwidth, nheight, 4, out);
    ipb_buffer_release(full);
    return rc;
  }
  *out = full;
  return IPB_OK;
}

int ipb_rotatecrop_run(ipb_ctx *ctx, const ipb_rotatecrop *op, ipb_buffer *in, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!op || !in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  auto passthrough = [&]() { ipb_buffer_retain(in); *out = in; return (int)IPB_OK; };
  if (rc_noop(op)) return passthrough();
  if (in->colors > 4) return fail(ctx, IPB_ERR_BAD_COLORS, "rotatecrop: %zu channels", in->colors);
  const float swidth = (float)in->width, sheight = (float)in->height;
  size_t nwidth, nheight;
  rc_calc_size(op, in->width, in->height, false, &nwidth, &nheight);
  const float fnwidth = (float)nwidth, fnheight = (float)nheight;
  float x = floorf(swidth * op->crop_left);
  if (x < 0.0f || x > swidth) return passthrough();  // rotatecrop.rs:49-52 (logs and returns the input)
  float y = floorf(sheight * op->crop_top);
  if (y < 0.0f || y > sheight) return passthrough();
  XformGeom g;
  rc_rotate_point_reverse(op, x, y, fnwidth, fnheight, swidth, sheight, g.tl);
  rc_rotate_point_reverse(op, x + fnwidth - 1.0f, y, fnwidth, fnheight, swidth, sheight, g.tr);
  rc_rotate_point_reverse(op, x, y + fnheight - 1.0f, fnwidth, fnheight, swidth, sheight, g.bl);
  g.width = in->width; g.height = in->height; g.nwidth = nwidth; g.nheight = nheight; g.components = in->colors;
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, nwidth, nheight, in->colors, in->monochrome, false, &b));
  cudaError_t e = launch_transform_f32(ctx->stream, g, nullptr, in->dptr, b->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "rotatecrop kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  *out = b;
  return IPB_OK;
}

static bool spline_counting_ok(const SplineDev &sp);
// to_lab, optionally with the basecurve applied in the same pass (sp != null, sp->n > 0)
static int tolab_launch(ipb_ctx *ctx, const ipb_tolab *op, const SplineDev *sp, ipb_buffer *in, ipb_buffer **out) {
  if (in->colors != 4) return fail(ctx, IPB_ERR_BAD_COLORS, "to_lab: expected 4 channels, got %zu", in->colors);
  ColorParams P;
  memset(&P, 0, sizeof(P));
  fill_tolab(&P, op, in->monochrome);
  if (sp) P.sp = *sp;
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, in->width, in->height, 3, in->monochrome, false, &b));
  const int curve = sp ? (spline_counting_ok(*sp) ? 2 : 1) : 0;
  cudaError_t e = launch_tolab(ctx->stream, P, ctx->lut_lab, ctx->cbrt_tab, curve, in->dptr, in->width * in->height, b->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "to_lab kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  *out = b;
  return IPB_OK;
}

int ipb_tolab_run(ipb_ctx *ctx, const ipb_tolab *op, ipb_buffer *in, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!op || !in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  return tolab_launch(ctx, op, nullptr, in, out);
}

int ipb_basecurve_run(ipb_ctx *ctx, const ipb_basecurve *op, ipb_buffer *in, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!op || !in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (op->npoints > IPB_MAX_CURVE_POINTS) return fail(ctx, IPB_ERR_UNSUPPORTED, "basecurve: more than %d points", IPB_MAX_CURVE_POINTS);
  SplineDev sp;
  if (!build_spline(op, &sp)) return fail(ctx, IPB_ERR_INVALID, "basecurve: degenerate curve (the reference panics)");
  if (sp.n == 0) {  // curves.rs:34-36
    ipb_buffer_retain(in);
    *out = in;
    return IPB_OK;
  }
  if (in->colors != 3) return fail(ctx, IPB_ERR_BAD_COLORS, "basecurve: expected 3 channels, got %zu", in->colors);
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, in->width, in->height, 3, in->monochrome, false, &b));
  cudaError_t e = launch_basecurve(ctx->stream, sp, in->dptr, in->width * in->height, b->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "basecurve kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  *out = b;
  return IPB_OK;
}

// from_lab, optionally with OpGamma applied in the same pass
static int fromlab_launch(ipb_ctx *ctx, bool with_gamma, ipb_buffer *in, ipb_buffer **out) {
  ColorParams P;
  memset(&P, 0, sizeof(P));
  memcpy(P.rgbm, tables().xyz_d65_33, sizeof(P.rgbm));
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, in->width, in->height, 3, in->monochrome, false, &b));
  cudaError_t e = launch_fromlab(ctx->stream, P, with_gamma ? ctx->lut_gamma : nullptr, in->dptr, in->width * in->height, b->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "from_lab kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  *out = b;
  return IPB_OK;
}

int ipb_fromlab_run(ipb_ctx *ctx, ipb_buffer *in, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (in->colors != 3) return fail(ctx, IPB_ERR_BAD_COLORS, "from_lab: expected 3 channels, got %zu", in->colors);
  return fromlab_launch(ctx, false, in, out);
}

int ipb_gamma_run(ipb_ctx *ctx, const ipb_settings *settings, ipb_buffer *in, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!settings || !in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (settings->linear) {  // gamma.rs:17-18
    ipb_buffer_retain(in);
    *out = in;
    return IPB_OK;
  }
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, in->width, in->height, in->colors, in->monochrome, false, &b));
  cudaError_t e = launch_gamma(ctx->stream, ctx->lut_gamma, in->dptr, in->width * in->height * in->colors, b->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "gamma kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  *out = b;
  return IPB_OK;
}

int ipb_transform_run(ipb_ctx *ctx, const ipb_transform *op, ipb_buffer *in, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!op || !in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  int f[3];
  orientation_flips(op, f);
  if (!f[0] && !f[1] && !f[2]) {  // transform.rs:68-69
    ipb_buffer_retain(in);
    *out = in;
    return IPB_OK;
  }
  if (in->colors != 3) return fail(ctx, IPB_ERR_BAD_COLORS, "transform: expected 3 channels, got %zu", in->colors);
  ipb_buffer *b;
  if (f[0]) IPB_TRY(new_buffer(ctx, in->height, in->width, 3, in->monochrome, false, &b));
  else IPB_TRY(new_buffer(ctx, in->width, in->height, 3, in->monochrome, false, &b));
  cudaError_t e = launch_rotate(ctx->stream, in->dptr, in->width, in->height, f[0], f[1], f[2], b->dptr);
  if (e != cudaSuccess) { ipb_buffer_release(b); return fail(ctx, IPB_ERR_CUDA, "rotate kernel: %s", cudaGetErrorString(e)); }
  ctx->launches++;
  *out = b;
  return IPB_OK;
}

// ---- host-only size negotiation

void ipb_gofloat_transform_forward(const ipb_gofloat *op, size_t w, size_t h, size_t *ow, size_t *oh) {
  size_t o[4];
  size_image(op, w, h, o);
  *ow = o[2];
  *oh = o[3];
}
void ipb_rotatecrop_transform_forward(ipb_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh) {
  if (op->has_output_size) { *ow = op->output_width; *oh = op->output_height; return; }  // rotatecrop.rs:67-69
  op->input_ratio = (float)w / (float)h;
  rc_calc_size(op, w, h, false, ow, oh);
}
void ipb_rotatecrop_transform_reverse(ipb_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh) {
  op->has_output_size = 1;
  op->output_width = w;
  op->output_height = h;
  rc_calc_size(op, w, h, true, ow, oh);
}
void ipb_rotatecrop_reset(ipb_rotatecrop *op) {
  op->input_ratio = 1.0f;
  op->has_output_size = 0;
}
void ipb_transform_transform_forward(const ipb_transform *op, size_t w, size_t h, size_t *ow, size_t *oh) {
  if (op->rotation == IPB_ROT_90 || op->rotation == IPB_ROT_270) { *ow = h; *oh = w; }
  else { *ow = w; *oh = h; }
}
void ipb_scaling_size(size_t w, size_t h, size_t maxw, size_t maxh, size_t *ow, size_t *oh) {
  float s;
  scaling_total(w, h, maxw, maxh, &s, ow, oh);
}
float ipb_calculate_scale(size_t w, size_t h, size_t maxw, size_t maxh) {
  float s;
  size_t a, b;
  scaling_total(w, h, maxw, maxh, &s, &a, &b);
  return s;
}

int ipb_spline_eval(ipb_ctx *ctx, const ipb_basecurve *op, const float *in, float *out, size_t n) {
  IPB_TRY(enter(ctx));
  if (!op || !in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (op->npoints > IPB_MAX_CURVE_POINTS) return fail(ctx, IPB_ERR_UNSUPPORTED, "too many curve points");
  ipb_basecurve raw = *op;
  raw.exposure = 0.0f;  // SplineFunc::new(&points) — no exposure scaling (curves.rs:53-55)
  SplineDev sp;
  memset(&sp, 0, sizeof(sp));
  if (raw.npoints == 0) {  // SplineFunc::new(&[]) is the identity line through (0,0) and (1,1)
    ipb_basecurve tmp = raw;
    tmp.exposure = 1.0f;  // defeat the pass-through shortcut of build_spline; y's are not scaled with no points
    if (!build_spline(&tmp, &sp)) return fail(ctx, IPB_ERR_INVALID, "degenerate curve");
  } else if (!build_spline(&raw, &sp)) {
    return fail(ctx, IPB_ERR_INVALID, "degenerate curve (the reference panics)");
  }
  float *d;
  IPB_CUDA(ctx, ipb_malloc_async(ctx, (void **)&d, 2 * (n ? n : 1) * sizeof(float)));
  IPB_CUDA(ctx, cudaMemcpyAsync(d, in, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  IPB_LAUNCH(ctx, launch_spline_eval(ctx->stream, sp, d, n, d + n));
  IPB_CUDA(ctx, cudaMemcpyAsync(out, d + n, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  IPB_CUDA(ctx, cudaFreeAsync(d, ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IPB_OK;
}

// ---- pack

extern "C++" {
template <typename T>
static int pack_impl(ipb_ctx *ctx, const ipb_buffer *in, T *dst, int dst_on_device) {
  IPB_TRY(enter(ctx));
  if (!in || !dst) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (in->colors != 3) return fail(ctx, IPB_ERR_BAD_COLORS, "pack: expected 3 channels, got %zu", in->colors);
  const size_t n = in->width * in->height * 3;
  T *d = dst;
  if (!dst_on_device) IPB_CUDA(ctx, ipb_malloc_async(ctx, (void **)&d, (n ? n : 1) * sizeof(T)));
  cudaError_t e = sizeof(T) == 1 ? launch_pack8(ctx->stream, in->dptr, n, (uint8_t *)d)
                                 : launch_pack16(ctx->stream, in->dptr, n, (uint16_t *)d);
  if (e != cudaSuccess) return fail(ctx, IPB_ERR_CUDA, "pack kernel: %s", cudaGetErrorString(e));
  ctx->launches++;
  if (!dst_on_device) {
    IPB_CUDA(ctx, cudaMemcpyAsync(dst, d, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    IPB_CUDA(ctx, cudaFreeAsync(d, ctx->stream));
    IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return IPB_OK;
}
}  // extern "C++"
int ipb_pack_8bit(ipb_ctx *ctx, const ipb_buffer *in, uint8_t *dst, int dst_on_device) {
  return pack_impl<uint8_t>(ctx, in, dst, dst_on_device);
}
int ipb_pack_16bit(ipb_ctx *ctx, const ipb_buffer *in, uint16_t *dst, int dst_on_device) {
  return pack_impl<uint16_t>(ctx, in, dst, dst_on_device);
}

extern "C++" {
template <typename T>
static int scale_srgb_impl(ipb_ctx *ctx, const T *src, size_t w, size_t h, size_t nw, size_t nh, T *dst, int on_device) {
  IPB_TRY(enter(ctx));
  if (!src || !dst) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (w == 0 || h == 0 || nw == 0 || nh == 0) return fail(ctx, IPB_ERR_INVALID, "empty image");
  const size_t nin = w * h * 3, nout = nw * nh * 3;
  const T *s = src;
  T *d = dst;
  T *tmp = nullptr;
  if (!on_device) {
    IPB_CUDA(ctx, ipb_malloc_async(ctx, (void **)&tmp, (nin + nout) * sizeof(T)));
    IPB_CUDA(ctx, cudaMemcpyAsync(tmp, src, nin * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    s = tmp;
    d = tmp + nin;
  }
  XformGeom g;
  g.tl[0] = 0; g.tl[1] = 0; g.tr[0] = (long)w - 1; g.tr[1] = 0; g.bl[0] = 0; g.bl[1] = (long)h - 1;
  g.width = w; g.height = h; g.nwidth = nw; g.nheight = nh; g.components = 3;
  cudaError_t e = sizeof(T) == 1 ? launch_transform_u8(ctx->stream, g, (const uint8_t *)s, (uint8_t *)d)
                                 : launch_transform_u16(ctx->stream, g, (const uint16_t *)s, (uint16_t *)d);
  if (e != cudaSuccess) return fail(ctx, IPB_ERR_CUDA, "scale_down_srgb kernel: %s", cudaGetErrorString(e));
  ctx->launches++;
  if (!on_device) {
    IPB_CUDA(ctx, cudaMemcpyAsync(dst, d, nout * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    IPB_CUDA(ctx, cudaFreeAsync(tmp, ctx->stream));
    IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return IPB_OK;
}
}  // extern "C++"
int ipb_scale_down_srgb(ipb_ctx *ctx, const uint8_t *src, size_t w, size_t h, size_t nw, size_t nh, uint8_t *dst,
                        int on_device) {
  return scale_srgb_impl<uint8_t>(ctx, src, w, h, nw, nh, dst, on_device);
}
int ipb_scale_down_srgb16(ipb_ctx *ctx, const uint16_t *src, size_t w, size_t h, size_t nw, size_t nh, uint16_t *dst,
                          int on_device) {
  return scale_srgb_impl<uint16_t>(ctx, src, w, h, nw, nh, dst, on_device);
}

// ------------------------------------------------------------------------------------------------ Pipeline

// ---- Lanczos-a resampler (extension; the reference only has the FIXME at scaling.rs:101-103)

extern "C++" {
namespace {
struct LzAxis {
  std::vector<int> start, count;
  std::vector<float> w;
  int ksize = 0;
};
double lz_kernel(double t, int a) {
  if (t < 0) t = -t;
  if (t >= (double)a) return 0.0;
  if (t == 0.0) return 1.0;
  const double pt = M_PI * t;
  return (sin(pt) / pt) * (sin(pt / (double)a) / (pt / (double)a));
}
// taps of one axis: centre (i + 0.5) * scale, support a * max(scale, 1), weights normalised in double, stored as f32
void lz_axis(size_t n_in, size_t n_out, int a, LzAxis *ax) {
  const double scale = (double)n_in / (double)n_out;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = (double)a * fscale;
  ax->ksize = (int)ceil(support) * 2 + 1;
  ax->start.assign(n_out, 0);
  ax->count.assign(n_out, 0);
  ax->w.assign(n_out * (size_t)ax->ksize, 0.0f);
  std::vector<double> tmp((size_t)ax->ksize);
  for (size_t i = 0; i < n_out; i++) {
    const double centre = ((double)i + 0.5) * scale;
    long xmin = (long)(centre - support + 0.5);
    long xmax = (long)(centre + support + 0.5);
    if (xmin < 0) xmin = 0;
    if (xmax > (long)n_in) xmax = (long)n_in;
    const long n = xmax - xmin;
    double sum = 0.0;
    for (long k = 0; k < n; k++) {
      tmp[(size_t)k] = lz_kernel(((double)(xmin + k) - centre + 0.5) / fscale, a);
      sum += tmp[(size_t)k];
    }
    for (long k = 0; k < n; k++)
      ax->w[i * (size_t)ax->ksize + (size_t)k] = (float)(sum != 0.0 ? tmp[(size_t)k] / sum : tmp[(size_t)k]);
    ax->start[i] = (int)xmin;
    ax->count[i] = (int)n;
  }
}
}  // namespace
}  // extern "C++"

// device-resident tap table of one axis, built once per (n_in, n_out, a) and context and kept (eight most recent)
static int lz_table(ipb_ctx *ctx, size_t n_in, size_t n_out, int a, const ipb_ctx::LzTab **out) {
  for (size_t i = 0; i < ctx->lz_tabs.size(); i++)
    if (ctx->lz_tabs[i].n_in == n_in && ctx->lz_tabs[i].n_out == n_out && ctx->lz_tabs[i].a == a) {
      ipb_ctx::LzTab t = std::move(ctx->lz_tabs[i]);
      ctx->lz_tabs.erase(ctx->lz_tabs.begin() + (long)i);
      ctx->lz_tabs.push_back(std::move(t));
      *out = &ctx->lz_tabs.back();
      return IPB_OK;
    }
  LzAxis ax;
  lz_axis(n_in, n_out, a, &ax);
  const size_t ibytes = ((2 * n_out * sizeof(int) + 15) / 16) * 16, wbytes = ax.w.size() * sizeof(float);
  char *dev = nullptr;
  IPB_CUDA(ctx, cudaMalloc((void **)&dev, ibytes + wbytes));
  std::vector<char> host(ibytes + wbytes);
  memcpy(host.data(), ax.start.data(), n_out * sizeof(int));
  memcpy(host.data() + n_out * sizeof(int), ax.count.data(), n_out * sizeof(int));
  memcpy(host.data() + ibytes, ax.w.data(), wbytes);
  cudaError_t e = cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // `host` is pageable and goes out of scope
  if (e != cudaSuccess) {
    cudaFree(dev);
    return fail(ctx, IPB_ERR_CUDA, "lanczos tables: %s", cudaGetErrorString(e));
  }
  if (ctx->lz_tabs.size() >= 8) {
    IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // a launch may still read the table we drop
    cudaFree(ctx->lz_tabs.front().start);
    ctx->lz_tabs.erase(ctx->lz_tabs.begin());
  }
  ipb_ctx::LzTab t;
  t.n_in = n_in; t.n_out = n_out; t.a = a; t.ksize = ax.ksize;
  t.start = (int *)dev; t.count = (int *)dev + n_out; t.w = (float *)(dev + ibytes);
  t.hstart = std::move(ax.start); t.hcount = std::move(ax.count);
  ctx->lz_tabs.push_back(std::move(t));
  *out = &ctx->lz_tabs.back();
  return IPB_OK;
}

int ipb_lanczos_resize(ipb_ctx *ctx, ipb_buffer *in, size_t nwidth, size_t nheight, int a, ipb_buffer **out) {
  IPB_TRY(enter(ctx));
  if (!in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (nwidth == 0 || nheight == 0 || a < 1 || a > 8 || in->width == 0 || in->height == 0)
    return fail(ctx, IPB_ERR_INVALID, "lanczos: sizes must be positive and 1 <= a <= 8");
  if (in->width >= (1u << 30) || in->height >= (1u << 30) || nwidth >= (1u << 30) || nheight >= (1u << 30))
    return fail(ctx, IPB_ERR_INVALID, "lanczos: frame too large");
  const size_t W = in->width, H = in->height, Cc = in->colors;
  const ipb_ctx::LzTab *tx = nullptr, *ty = nullptr;
  IPB_TRY(lz_table(ctx, W, nwidth, a, &tx));
  // copy what we need of the x table: fetching the y table may reorder / evict entries of the small cache
  const int *sx = tx->start, *cx = tx->count;
  const float *wx = tx->w;
  const int kx = tx->ksize;
  // widest input span a 256-column tile of the horizontal pass stages in shared memory
  size_t max_span = 0;
  for (size_t x0 = 0; x0 < nwidth; x0 += 256) {
    const size_t x1 = x0 + 256 < nwidth ? x0 + 256 : nwidth;
    const size_t span = (size_t)(tx->hstart[x1 - 1] + tx->hcount[x1 - 1] - tx->hstart[x0]) * Cc;
    if (span > max_span) max_span = span;
  }
  if ((max_span + 8) * sizeof(float) > 200 * 1024)
    return fail(ctx, IPB_ERR_UNSUPPORTED, "lanczos: a 256-column tile spans %zu floats of a source row (scale too large)", max_span);
  IPB_TRY(lz_table(ctx, H, nheight, a, &ty));
  ipb_buffer *o;
  IPB_TRY(new_buffer(ctx, nwidth, nheight, Cc, in->monochrome, false, &o));
  float *mid = nullptr;
  cudaError_t e = ipb_malloc_async(ctx, (void **)&mid, H * nwidth * Cc * sizeof(float));
  if (e != cudaSuccess) { ipb_buffer_release(o); return fail(ctx, IPB_ERR_NOMEM, "lanczos: %s", cudaGetErrorString(e)); }
  e = launch_lanczos(ctx->stream, in->dptr, W, H, Cc, nwidth, nheight, sx, cx, wx, kx, max_span, ty->start, ty->count, ty->w,
                     ty->ksize, mid, o->dptr);
  cudaFreeAsync(mid, ctx->stream);
  if (e != cudaSuccess) { ipb_buffer_release(o); return fail(ctx, IPB_ERR_CUDA, "lanczos: %s", cudaGetErrorString(e)); }
  ctx->launches += 2;
  *out = o;
  return IPB_OK;
}

void ipb_ops_default(ipb_ops *ops, const ipb_source *image) {
  const HostTables &T = tables();
  memset(ops, 0, sizeof(*ops));
  ops->rotatecrop.input_ratio = 1.0f;  // rotatecrop.rs:27-36
  const bool raw = image->kind == IPB_SRC_RAW_U16 || image->kind == IPB_SRC_RAW_F32;
  if (raw) {
    ops->gofloat.is_cfa = 1;  // filled from metadata by the caller (gofloat.rs:20-31)
    ops->basecurve.npoints = 1;  // curves.rs:14-20
    ops->basecurve.points[0][0] = 0.50f;
    ops->basecurve.points[0][1] = 0.60f;
    for (int i = 0; i < 4; i++) ops->tolab.wb_coeffs[i] = 1.0f;
  } else {
    memcpy(ops->tolab.cam_to_xyz, T.srgb_d65_43, sizeof(T.srgb_d65_43));  // colorspaces.rs:48-55
    memcpy(ops->tolab.cam_to_xyz_normalized, T.srgb_d65_43, sizeof(T.srgb_d65_43));
    memcpy(ops->tolab.xyz_to_cam, T.xyz_d65_34, sizeof(T.xyz_d65_34));
    ops->tolab.wb_coeffs[0] = 1.0f; ops->tolab.wb_coeffs[1] = 1.0f; ops->tolab.wb_coeffs[2] = 1.0f;
    ops->tolab.wb_coeffs[3] = 0.0f;
  }
}

int ipb_pipeline_create(ipb_ctx *ctx, const ipb_source *image, const ipb_ops *ops, ipb_pipeline **out) {
  IPB_TRY(enter(ctx));
  if (!image || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  ipb_pipeline *p = new (std::nothrow) ipb_pipeline();
  if (!p) return fail(ctx, IPB_ERR_NOMEM, "out of host memory");
  p->ctx = ctx;
  p->image = *image;
  if (ops) p->ops = *ops; else ipb_ops_default(&p->ops, image);
  memset(&p->settings, 0, sizeof(p->settings));
  p->settings.use_fastpath = 1;  // pipeline.rs:121-130
  *out = p;
  return IPB_OK;
}

void ipb_pipeline_destroy(ipb_pipeline *p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  if (p->stage_in) cudaFree(p->stage_in);
  if (p->stage_out) cudaFree(p->stage_out);
  delete p;
}

ipb_ops *ipb_pipeline_ops(ipb_pipeline *p) { return p ? &p->ops : nullptr; }
ipb_settings *ipb_pipeline_settings(ipb_pipeline *p) { return p ? &p->settings : nullptr; }
int ipb_pipeline_set_source(ipb_pipeline *p, const ipb_source *image) {
  if (!p || !image) return IPB_ERR_INVALID;
  p->image = *image;
  p->has_stripe = false;
  p->source_gen++;
  return IPB_OK;
}
int ipb_pipeline_set_fused(ipb_pipeline *p, int fused) {
  if (!p) return IPB_ERR_INVALID;
  p->fused = fused != 0;
  return IPB_OK;
}

// pipeline.rs:313-338: reset, forward size walk, clamp to maxwidth/maxheight, reverse walk
static void negotiate(ipb_pipeline *p, size_t *fw, size_t *fh) {
  ipb_rotatecrop_reset(&p->ops.rotatecrop);
  size_t width = p->image.width, height = p->image.height, w, h;
  ipb_gofloat_transform_forward(&p->ops.gofloat, width, height, &w, &h); width = w; height = h;
  ipb_rotatecrop_transform_forward(&p->ops.rotatecrop, width, height, &w, &h); width = w; height = h;
  ipb_transform_transform_forward(&p->ops.transform, width, height, &w, &h); width = w; height = h;
  ipb_scaling_size(width, height, p->settings.maxwidth, p->settings.maxheight, &w, &h); width = w; height = h;
  if (fw) *fw = width;
  if (fh) *fh = height;
  ipb_transform_transform_forward(&p->ops.transform, width, height, &w, &h); width = w; height = h;  // transform.rs:82-84
  ipb_rotatecrop_transform_reverse(&p->ops.rotatecrop, width, height, &w, &h); width = w; height = h;
  p->settings.demosaic_width = width;
  p->settings.demosaic_height = height;
}

int ipb_pipeline_output_size(ipb_pipeline *p, size_t *width, size_t *height) {
  if (!p) return IPB_ERR_INVALID;
  if (p->image.width < 10 || p->image.height < 10) return fail(p->ctx, IPB_ERR_INVALID, "source smaller than 10x10");
  negotiate(p, width, height);
  return IPB_OK;
}

// What the fused kernels can take over: a u16 CFA source through gofloat's CFA branch, a pass-through
// rotatecrop, and either demosaic branch that starts from the 1-channel buffer without an intermediate
// full-size image (demosaic.rs:47-50 scaled, :51-60 with scale <= 1 full).
enum FusedMode { kNotFused = 0, kFusedFull = 1, kFusedScaled = 2 };
struct FusedPlan {
  FusedMode mode = kNotFused;
  CfaDev cfa;
  size_t crop_x = 0, crop_y = 0, width = 0, height = 0;  // cropped frame
  size_t out_width = 0, out_height = 0;                  // demosaic output == fused output
};


// The fused kernels divide by the colour chain's constants with a 3-instruction reciprocal form that equals IEEE
// division only while the dividend and the quotient are finite and normal (or zero) — tools/verify_constdiv.c.
// These bounds keep every dividend of the chain in that range for any u16 sample: levels within the u16 range and
// at least one code value apart, white-balance multipliers and matrix entries zero or between 2^-20 and 64 in
// magnitude.  Real camera metadata is far inside them; anything else runs op by op with IEEE division.
// Can the spline be evaluated by counting the knots at or below the value (k_fused_*, k_spec8, k_tolab<2>)?  Knots must be
// finite and increase strictly; the coefficients must be finite: on a knot the counting form returns y + 0 * c, which is
// NaN for an infinite c (tiny knot spacing overflows 1/dx) where the reference's early return gives y; and the end values
// must not be -0.0 (they are returned through y + 0 * d sums).
static bool spline_counting_ok(const SplineDev &sp) {
  for (int i = 0; i < sp.n; i++)
    if (!std::isfinite(sp.x[i]) || !std::isfinite(sp.y[i])) return false;
  for (int i = 0; i + 1 < sp.n; i++)
    if (!(sp.x[i] < sp.x[i + 1])) return false;
  for (int i = 0; i < sp.nseg; i++)
    if (!std::isfinite(sp.c1[i]) || !std::isfinite(sp.c2[i]) || !std::isfinite(sp.c3[i])) return false;
  if (sp.n > 0 && (std::signbit(sp.y_first) && sp.y_first == 0.0f)) return false;
  if (sp.n > 0 && (std::signbit(sp.y_last) && sp.y_last == 0.0f)) return false;
  return true;
}

static bool fused_params_bounded(const ipb_pipeline *p) {
  const ipb_gofloat &g = p->ops.gofloat;
  const float black = g.blacklevels[0], range = g.whitelevels[0] - g.blacklevels[0];
  // (a negative range — white below black — would turn exact zeros into -0.0, which the kernels' "+0.0 is a no-op"
  // sums do not reproduce; such metadata runs op by op)
  if (!std::isfinite(black) || !std::isfinite(range) || fabsf(black) > 65535.0f || !(range >= 1.0f) || range > 131072.0f)
    return false;
  float mul[4];
  normalize_wbs(p->ops.tolab.wb_coeffs, mul);
  auto ok = [](float v) { return std::isfinite(v) && (v == 0.0f || (fabsf(v) >= 9.5367431640625e-07f && fabsf(v) <= 64.0f)); };
  for (int i = 0; i < 4; i++)
    if (!ok(mul[i])) return false;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++)
      if (!ok(p->ops.tolab.cam_to_xyz_normalized[i][j])) return false;
  const ipb_basecurve &c = p->ops.basecurve;
  if (!std::isfinite(c.exposure) || fabsf(c.exposure) > 16.0f) return false;
  for (size_t i = 0; i < c.npoints && i < IPB_MAX_CURVE_POINTS; i++)
    if (!std::isfinite(c.points[i][0]) || !std::isfinite(c.points[i][1]) || fabsf(c.points[i][0]) > 1024.0f ||
        fabsf(c.points[i][1]) > 1024.0f)
      return false;
  // the fused spline evaluation counts the knots at or below the value: knots must increase strictly, and the end
  // values must not be -0.0 (they are returned through y + 0*d sums)
  SplineDev sp;
  if (!build_spline(&c, &sp)) return false;
  return spline_counting_ok(sp);
}

static int plan_fused(ipb_pipeline *p, FusedPlan *plan) {
  plan->mode = kNotFused;
  if (!p->fused) return IPB_OK;
  const ipb_source &img = p->image;
  if (img.kind != IPB_SRC_RAW_U16 || img.cpp != 1 || !p->ops.gofloat.is_cfa) return IPB_OK;
  if (!rc_noop(&p->ops.rotatecrop)) return IPB_OK;
  if (parse_cfa(p->ops.demosaic.cfa, &plan->cfa) != IPB_OK || plan->cfa.width == 0) return IPB_OK;
  if (p->ops.basecurve.npoints > IPB_MAX_CURVE_POINTS) return IPB_OK;
  if (!fused_params_bounded(p)) return IPB_OK;
  if (img.width >= (1u << 30) || img.height >= (1u << 30)) return IPB_OK;
  size_t xywh[4];
  size_image(&p->ops.gofloat, img.width, img.height, xywh);
  plan->crop_x = xywh[0]; plan->crop_y = xywh[1]; plan->width = xywh[2]; plan->height = xywh[3];
  float scale;
  size_t sw, sh;
  scaling_total(plan->width, plan->height, p->settings.demosaic_width, p->settings.demosaic_height, &scale, &sw, &sh);
  if (scale >= cfa_minscale(plan->cfa)) {
    plan->mode = kFusedScaled;
    plan->out_width = p->settings.demosaic_width;
    plan->out_height = p->settings.demosaic_height;
    if (plan->out_width < 2 || plan->out_height < 2) plan->mode = kNotFused;
  } else if (scale <= 1.0f) {
    plan->mode = kFusedFull;
    plan->out_width = plan->width;
    plan->out_height = plan->height;
  }
  return IPB_OK;
}

static int fill_color_params(ipb_pipeline *p, const FusedPlan &plan, ColorParams *P) {
  memset(P, 0, sizeof(*P));
  fill_tolab(P, &p->ops.tolab, 0);
  // the E channel is identically zero when the pattern has no colour 3 and its matrix column is finite
  bool has_e = false;
  for (int i = 0; i < 48 * 48; i++) has_e |= plan.cfa.pat[i] == 3;
  P->use_e = (has_e || !std::isfinite(P->cm[3]) || !std::isfinite(P->cm[7]) || !std::isfinite(P->cm[11]) ||
              !std::isfinite(P->mul[3])) ? 1 : 0;
  P->linear = p->settings.linear ? 1 : 0;
  P->one = 1.0f;
  P->mone = -1.0f;
  if (!build_spline(&p->ops.basecurve, &P->sp)) return fail(p->ctx, IPB_ERR_INVALID, "basecurve: degenerate curve");
  return IPB_OK;
}

// source rows (un-cropped sensor coordinates) needed for output rows [r0, r1) of the fused path
static void fused_src_rows(const ipb_pipeline *p, const FusedPlan &plan, size_t r0, size_t r1, size_t *s0, size_t *s1) {
  size_t a, b;
  if (plan.mode == kFusedFull) {
    a = r0 > 0 ? r0 - 1 : 0;
    b = r1 + 1 < plan.height ? r1 + 1 : plan.height;
  } else {
    // scaling.rs:72,79-80,86-87 with topleft (0,0): from_y = floor(skip_y_y * row), to_y = floor(skip_y_y * (row+1))
    const float skip_y = ((float)((long)plan.height - 1) - 0.0f) / (float)(plan.out_height - 1);
    auto clampy = [&](float f) { size_t v = f2usize(floorf(f)); return v < plan.height - 1 ? v : plan.height - 1; };
    a = clampy(0.0f + skip_y * (float)r0);
    b = clampy(0.0f + skip_y * (float)r1) + 1;  // to_y of the last row r1-1 uses (row+1) == r1
  }
  *s0 = a + plan.crop_y;
  *s1 = b + plan.crop_y;
  (void)p;
}

static int ensure_stage(ipb_ctx *ctx, void **buf, size_t *have, size_t need) {
  if (*have >= need) return IPB_OK;
  if (*buf) {
    IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    IPB_CUDA(ctx, cudaFree(*buf));
    *buf = nullptr;
    *have = 0;
  }
  IPB_CUDA(ctx, cudaMalloc(buf, need));
  *have = need;
  return IPB_OK;
}

// XYZ_LAB_TRANSFORM's analytic branch above 1.0 is v.cbrt() (color_conversions.rs:123): the host libm's cbrtf, like
// the 8193-entry tables.  The fused full-resolution kernel reads it from a table of every float in (1.0, 1.5] —
// 2^22 entries, 16 MB, built once per context with the same libm call (about 60 ms) — instead of restating glibc's
// double-precision algorithm per value; ratios beyond 1.5 still take that restatement (lab_f_slow).
static int ensure_cbrt_table(ipb_ctx *ctx) {
  if (ctx->cbrt_tab) return IPB_OK;
  const size_t n = (size_t)1 << 22;
  std::vector<float> host(n);
  for (size_t i = 0; i < n; i++) {
    const uint32_t bits = 0x3f800001u + (uint32_t)i;
    float v;
    memcpy(&v, &bits, 4);
    host[i] = cbrtf(v);
  }
  float *d = nullptr;
  IPB_CUDA(ctx, cudaMalloc((void **)&d, n * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(d, host.data(), n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(ctx, IPB_ERR_CUDA, "cube-root table upload: %s", cudaGetErrorString(e));
  }
  ctx->cbrt_tab = d;
  return IPB_OK;
}

// Tables and constants of the speculative kernel for this parameter set, rebuilt (and uploaded in stream order) only
// when the colour parameters, the level mapping or the forced bound change.  *use = false: these parameters stay on
// k_fused_full.
static int ensure_spec_tables(ipb_ctx *ctx, const ColorParams &P, float black, float range, bool *use) {
  *use = false;
  if (!ctx->spec_ok) return IPB_OK;
  std::vector<unsigned char> key(sizeof(ColorParams) + 3 * sizeof(float));
  memcpy(key.data(), &P, sizeof(ColorParams));
  memcpy(key.data() + sizeof(ColorParams), &black, 4);
  memcpy(key.data() + sizeof(ColorParams) + 4, &range, 4);
  memcpy(key.data() + sizeof(ColorParams) + 8, &ctx->spec_delta_override, 4);
  if (ctx->spec_tab_state != 0 && key == ctx->spec_key) {
    *use = ctx->spec_tab_state > 0;
    return IPB_OK;
  }
  ctx->spec_key = key;
  ctx->spec_tab_state = -1;
  std::vector<float> thr;
  for (const Gamma8Entry &g : g_gamma8)
    if (g.thr <= 1.0f) thr.push_back(g.thr);
  std::vector<uint32_t> g8a;
  std::vector<float2> stab;
  SpecTables T{};
  float dl[4] = {0, 0, 0, 0};
  if (!spec_build(P, black, range, ctx->mufu_cbrt_err, ctx->spec_delta_override, thr, &g8a, &stab, &T.consts, dl))
    return IPB_OK;
  // the previous upload may still be in flight from the pinned staging buffer
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned char *st = (unsigned char *)ctx->spec_stage;
  memcpy(st, g8a.data(), kSpecG8Entries * sizeof(uint32_t));
  memcpy(st + kSpecG8Entries * sizeof(uint32_t), stab.data(), kSpecSTabEntries * sizeof(float2));
  IPB_CUDA(ctx, cudaMemcpyAsync(ctx->spec_g8a, st, kSpecG8Entries * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  IPB_CUDA(ctx, cudaMemcpyAsync(ctx->spec_stab, st + kSpecG8Entries * sizeof(uint32_t), kSpecSTabEntries * sizeof(float2),
                                cudaMemcpyHostToDevice, ctx->stream));
  T.delta = dl[0];
  for (int ch = 0; ch < 3; ch++) T.delta_ch[ch] = dl[1 + ch];
  T.g8a = ctx->spec_g8a;
  T.stab = ctx->spec_stab;
  T.thr8 = ctx->spec_thr;
  T.stats = ctx->spec_stats;
  ctx->spec_tab = T;
  ctx->spec_tab_state = 1;
  *use = true;
  return IPB_OK;
}

// Launch the fused kernel for output rows [r0, r1) into `out` (device, row r0 first).  `raw_dev` holds the
// un-cropped source rows [have0, have0 + have_rows) on the device.
static int launch_fused_rows(ipb_pipeline *p, const FusedPlan &plan, const ColorParams &P, int out_kind, size_t r0,
                             size_t r1, void *out, const uint16_t *raw_dev, size_t have0, size_t have_rows) {
  ipb_ctx *ctx = p->ctx;
  size_t need0, need1;
  fused_src_rows(p, plan, r0, r1, &need0, &need1);
  if (need0 < have0 || need1 > have0 + have_rows)
    return fail(ctx, IPB_ERR_INVALID, "stripe holds source rows [%zu,%zu) but output rows [%zu,%zu) need [%zu,%zu)", have0,
                have0 + have_rows, r0, r1, need0, need1);
  const size_t pitch = p->image.width;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.raw = raw_dev + (need0 - have0) * pitch;
  a.raw_pitch = pitch;
  a.src_row0 = need0;
  a.src_rows = need1 - need0;
  a.crop_x = plan.crop_x; a.crop_y = plan.crop_y;
  a.width = plan.width; a.height = plan.height;
  a.out_row0 = r0; a.out_row1 = r1;
  a.out_width = plan.out_width; a.out_height = plan.out_height;
  a.out = out;
  a.out_kind = out_kind;
  a.black = p->ops.gofloat.blacklevels[0];
  a.range = p->ops.gofloat.whitelevels[0] - a.black;  // gofloat.rs:86-89
  a.range_rc = 1.0f / a.range;
  if (!p->rc_cached || memcmp(&p->rc_black, &a.black, 4) != 0 || memcmp(&p->rc_range, &a.range, 4) != 0) {
    p->rc_exact = golevel_rc_exact(a.black, a.range, a.range_rc) ? 1 : 0;
    p->rc_black = a.black;
    p->rc_range = a.range;
    p->rc_cached = true;
  }
  a.exact_rc = p->rc_exact;
  a.lut_lab = ctx->lut_lab;
  a.lut_gamma = ctx->lut_gamma;
  a.lut_gamma8 = ctx->lut_gamma8;
  const bool scaled_spec = plan.mode == kFusedScaled && p->spec && out_kind == kOutU8;
  if (plan.mode == kFusedFull || scaled_spec) IPB_TRY(ensure_cbrt_table(ctx));
  a.cbrt_tab = ctx->cbrt_tab;
  a.use_tma = p->use_tma;
  // 8-bit output of a full-resolution RGB Bayer frame: the speculative kernel (byte-identical, about a third of the
  // instructions), when its preconditions hold
  if (p->batch_n > 1) { a.batch_n = p->batch_n; a.batch_src_rows = p->batch_src_rows; a.batch_out_bytes = p->batch_out_bytes; }
  if (plan.mode == kFusedFull && p->spec && a.exact_rc && spec_supported(a, plan.cfa, P)) {
    bool use = false;
    IPB_TRY(ensure_spec_tables(ctx, P, a.black, a.range, &use));
    if (use) {
      cudaError_t es = launch_fused_spec8(ctx->stream, a, plan.cfa, P, ctx->spec_tab, ctx->sm_count, ctx->spec_threads);
      if (es != cudaSuccess) return fail(ctx, IPB_ERR_CUDA, "speculative kernel: %s %s", cudaGetErrorString(es), spec_last_error());
      ctx->launches++;
      if (p->batch_n > 1) p->batch_done = true;
      return IPB_OK;
    }
  }
  if (p->batch_n > 1) return IPB_ERR_UNSUPPORTED;   // only k_spec8 takes a batch: the caller runs the frames one by one
  // ... and of a down-scaled one: the same chain behind scaled_demosaic's window phase
  if (scaled_spec && spec_scaled_supported(a, plan.cfa, P)) {
    bool use = false;
    IPB_TRY(ensure_spec_tables(ctx, P, a.black, a.range, &use));
    if (use) {
      cudaError_t es = launch_scaled_spec8(ctx->stream, a, plan.cfa, P, ctx->spec_tab, ctx->sm_count);
      if (es != cudaSuccess) return fail(ctx, IPB_ERR_CUDA, "speculative scaled kernel: %s %s", cudaGetErrorString(es), spec_last_error());
      ctx->launches++;
      return IPB_OK;
    }
  }
  cudaError_t e = plan.mode == kFusedFull ? launch_fused_full(ctx->stream, a, plan.cfa, P, ctx->sm_count)
                                          : launch_fused_scaled(ctx->stream, a, plan.cfa, P, ctx->sm_count);
  if (e != cudaSuccess) return fail(ctx, IPB_ERR_CUDA, "fused kernel: %s %s", cudaGetErrorString(e), fused_last_error());
  ctx->launches++;
  return IPB_OK;
}

// Output rows [r0, r1) through the fused kernel in one launch; a host-resident source is staged first.
static int run_fused(ipb_pipeline *p, const FusedPlan &plan, int out_kind, size_t r0, size_t r1, void *out) {
  ipb_ctx *ctx = p->ctx;
  ColorParams P;
  IPB_TRY(fill_color_params(p, plan, &P));
  // which source rows do we have?
  const ipb_source &src = p->has_stripe ? p->stripe_rows : p->image;
  const size_t have0 = p->has_stripe ? p->stripe.src_row0 : 0;
  const size_t have1 = have0 + src.height;
  const uint16_t *raw = (const uint16_t *)src.data;
  if (src.on_device) return launch_fused_rows(p, plan, P, out_kind, r0, r1, out, raw, have0, src.height);
  size_t need0, need1;
  fused_src_rows(p, plan, r0, r1, &need0, &need1);
  if (need0 < have0 || need1 > have1)
    return fail(ctx, IPB_ERR_INVALID, "stripe holds source rows [%zu,%zu) but output rows [%zu,%zu) need [%zu,%zu)", have0,
                have1, r0, r1, need0, need1);
  // copy just the rows this launch needs
  const size_t bytes = (need1 - need0) * src.width * sizeof(uint16_t);
  IPB_TRY(ensure_stage(ctx, &p->stage_in, &p->stage_in_bytes, bytes));
  IPB_CUDA(ctx, cudaMemcpyAsync(p->stage_in, raw + (need0 - have0) * src.width, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return launch_fused_rows(p, plan, P, out_kind, r0, r1, out, (const uint16_t *)p->stage_in, need0, need1 - need0);
}

static int ensure_copy_streams(ipb_ctx *ctx, size_t nevents) {
  if (!ctx->copy_in) IPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
  if (!ctx->copy_out) IPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
  while (ctx->events.size() < nevents) {
    cudaEvent_t e;
    IPB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->events.push_back(e);
  }
  return IPB_OK;
}

// Host-resident source and/or destination: the frame is cut into bands of output rows and the three legs of every
// band — H2D of the source rows it adds, the fused kernel, D2H of its output rows — run on three streams, so that
// the copies of neighbouring bands overlap each other (PCIe is full duplex) and the kernel.  The kernel stays on the
// context's stream.  Output rows [r0, r1); `dst` receives row r0 first.  Results are those of one whole-frame launch:
// every launch works in full-frame coordinates.
static int run_fused_banded(ipb_pipeline *p, const FusedPlan &plan, int out_kind, size_t elem_size, size_t r0, size_t r1,
                            void *dst, int dst_on_device) {
  ipb_ctx *ctx = p->ctx;
  const ipb_source &src = p->has_stripe ? p->stripe_rows : p->image;
  const size_t have0 = p->has_stripe ? p->stripe.src_row0 : 0;
  const size_t have1 = have0 + src.height;
  const bool src_host = !src.on_device, dst_host = !dst_on_device;
  const size_t row_out_bytes = plan.out_width * 3 * elem_size;
  const size_t row_in_bytes = src.width * sizeof(uint16_t);
  // bands of about 8 MB of traffic, at most 16, a multiple of 32 output rows (the full-resolution kernel's tile height)
  const size_t rows = r1 - r0;
  size_t need0, need1;
  fused_src_rows(p, plan, r0, r1, &need0, &need1);
  if (need0 < have0 || need1 > have1)
    return fail(ctx, IPB_ERR_INVALID, "stripe holds source rows [%zu,%zu) but output rows [%zu,%zu) need [%zu,%zu)", have0,
                have1, r0, r1, need0, need1);
  const size_t traffic = (src_host ? (need1 - need0) * row_in_bytes : 0) + (dst_host ? rows * row_out_bytes : 0);
  const size_t band_mb = p->band_mb > 0 ? (size_t)p->band_mb : ((size_t)1 << 40);  // 0: one band
  size_t nbands = traffic / (band_mb << 20);
  nbands = nbands < 1 ? 1 : (nbands > 64 ? 64 : nbands);
  size_t band_rows = ((rows + nbands - 1) / nbands + 31) / 32 * 32;
  nbands = (rows + band_rows - 1) / band_rows;
  if (nbands <= 1 || (!src_host && !dst_host)) {
    void *d = dst;
    if (dst_host) {
      IPB_TRY(ensure_stage(ctx, &p->stage_out, &p->stage_out_bytes, rows * row_out_bytes));
      d = p->stage_out;
    }
    IPB_TRY(run_fused(p, plan, out_kind, r0, r1, d));
    if (dst_host) IPB_CUDA(ctx, cudaMemcpyAsync(dst, d, rows * row_out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return IPB_OK;
  }
  ColorParams P;
  IPB_TRY(fill_color_params(p, plan, &P));
  IPB_TRY(ensure_copy_streams(ctx, 2 * nbands + 1));
  const uint16_t *raw_dev = (const uint16_t *)src.data;
  size_t dev0 = have0, dev_rows = src.height;
  if (src_host) {
    IPB_TRY(ensure_stage(ctx, &p->stage_in, &p->stage_in_bytes, (need1 - need0) * row_in_bytes));
    raw_dev = (const uint16_t *)p->stage_in;
    dev0 = need0;
    dev_rows = need1 - need0;
  }
  uint8_t *out_dev = (uint8_t *)dst;
  if (dst_host) {
    IPB_TRY(ensure_stage(ctx, &p->stage_out, &p->stage_out_bytes, rows * row_out_bytes));
    out_dev = (uint8_t *)p->stage_out;
  }
  // the copy streams start after whatever the context's stream has queued so far (it may still use the staging buffers)
  cudaEvent_t e0 = ctx->events[2 * nbands];
  IPB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  if (src_host) IPB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, e0, 0));
  if (dst_host) IPB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, e0, 0));
  // IPB_TRACE=1: time every leg with CUDA events and print the band timeline to stderr (debugging aid, slow)
  static const bool trace = getenv("IPB_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tev.push_back(e);
  };
  mark(ctx->stream);
  size_t uploaded = need0;  // source rows [need0, uploaded) are on their way
  for (size_t b = 0; b < nbands; b++) {
    const size_t b0 = r0 + b * band_rows, b1 = b0 + band_rows < r1 ? b0 + band_rows : r1;
    if (src_host) {
      size_t s0, s1;
      fused_src_rows(p, plan, b0, b1, &s0, &s1);
      if (s1 > uploaded) {
        IPB_CUDA(ctx, cudaMemcpyAsync((uint8_t *)p->stage_in + (uploaded - need0) * row_in_bytes,
                                      (const uint8_t *)src.data + (uploaded - have0) * row_in_bytes,
                                      (s1 - uploaded) * row_in_bytes, cudaMemcpyHostToDevice, ctx->copy_in));
        uploaded = s1;
      }
      IPB_CUDA(ctx, cudaEventRecord(ctx->events[2 * b], ctx->copy_in));
      IPB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->events[2 * b], 0));
    }
    mark(ctx->copy_in);
    uint8_t *band_out = out_dev + (b0 - r0) * row_out_bytes;
    mark(ctx->stream);
    IPB_TRY(launch_fused_rows(p, plan, P, out_kind, b0, b1, band_out, raw_dev, dev0, dev_rows));
    mark(ctx->stream);
    if (dst_host) {
      IPB_CUDA(ctx, cudaEventRecord(ctx->events[2 * b + 1], ctx->stream));
      IPB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->events[2 * b + 1], 0));
      IPB_CUDA(ctx, cudaMemcpyAsync((uint8_t *)dst + (b0 - r0) * row_out_bytes, band_out, (b1 - b0) * row_out_bytes,
                                    cudaMemcpyDeviceToHost, ctx->copy_out));
    }
    mark(ctx->copy_out);
  }
  if (trace) {
    cudaDeviceSynchronize();
    fprintf(stderr, "band  h2d_done  k_start  k_end  d2h_done   (ms after the call started)\n");
    for (size_t b = 0; b < nbands; b++) {
      float t[4];
      for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], tev[0], tev[1 + 4 * b + k]);
      fprintf(stderr, "%4zu  %8.3f %8.3f %6.3f %9.3f\n", b, t[0], t[1], t[2], t[3]);
    }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
  }
  if (dst_host) {  // the context's stream is "done" only when the last band has landed on the host
    IPB_CUDA(ctx, cudaEventRecord(e0, ctx->copy_out));
    IPB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e0, 0));
  }
  return IPB_OK;
}

// op-by-op Pipeline::run (pipeline.rs:364-372 with cache == None), starting from the source
static int run_unfused(ipb_pipeline *p, ipb_buffer **out) {
  ipb_ctx *ctx = p->ctx;
  ipb_buffer *cur = nullptr, *next = nullptr;
  IPB_TRY(ipb_gofloat_run(ctx, &p->ops.gofloat, &p->image, &cur));
#define STEP(call)                                  \
  do {                                              \
    int rc_ = (call);                               \
    ipb_buffer_release(cur);                        \
    if (rc_ != IPB_OK) return rc_;                  \
    cur = next;                                     \
  } while (0)
  STEP(ipb_demosaic_run(ctx, &p->ops.demosaic, &p->settings, cur, &next));
  STEP(ipb_rotatecrop_run(ctx, &p->ops.rotatecrop, cur, &next));
  // Nobody sees the buffers between to_lab and basecurve or between from_lab and gamma here (Pipeline::run without a
  // cache), so each pair runs as one pass with the same per-pixel arithmetic: two buffers fewer through HBM.  Anything
  // unusual (a curve the per-op entry point rejects or passes through, wrong channel counts, linear output) takes the
  // separate ops, which report it exactly as before.
  {
    SplineDev sp;
    const bool pair = p->ops.basecurve.npoints <= IPB_MAX_CURVE_POINTS && build_spline(&p->ops.basecurve, &sp) && sp.n > 0 &&
                      cur->colors == 4;
    if (pair) {
      STEP(tolab_launch(ctx, &p->ops.tolab, &sp, cur, &next));
    } else {
      STEP(ipb_tolab_run(ctx, &p->ops.tolab, cur, &next));
      STEP(ipb_basecurve_run(ctx, &p->ops.basecurve, cur, &next));
    }
  }
  if (!p->settings.linear && cur->colors == 3) {
    STEP(fromlab_launch(ctx, true, cur, &next));
  } else {
    STEP(ipb_fromlab_run(ctx, cur, &next));
    STEP(ipb_gamma_run(ctx, &p->settings, cur, &next));
  }
  STEP(ipb_transform_run(ctx, &p->ops.transform, cur, &next));
#undef STEP
  *out = cur;
  return IPB_OK;
}

int ipb_pipeline_run(ipb_pipeline *p, ipb_buffer **out) {
  if (!p || !out) return IPB_ERR_INVALID;
  ipb_ctx *ctx = p->ctx;
  IPB_TRY(enter(ctx));
  if (p->image.width < 10 || p->image.height < 10) return fail(ctx, IPB_ERR_INVALID, "source smaller than 10x10");
  if (p->has_stripe) return fail(ctx, IPB_ERR_UNSUPPORTED, "pipeline_run on a stripe source: use output_8bit_stripe");
  negotiate(p, nullptr, nullptr);
  FusedPlan plan;
  IPB_TRY(plan_fused(p, &plan));
  if (plan.mode == kNotFused) return run_unfused(p, out);
  ipb_buffer *b;
  IPB_TRY(new_buffer(ctx, plan.out_width, plan.out_height, 3, 0, false, &b));
  int rc = run_fused(p, plan, kOutF32, 0, plan.out_height, b->dptr);
  if (rc != IPB_OK) { ipb_buffer_release(b); return rc; }
  ipb_buffer *t;
  rc = ipb_transform_run(ctx, &p->ops.transform, b, &t);
  ipb_buffer_release(b);
  if (rc != IPB_OK) return rc;
  *out = t;
  return IPB_OK;
}

// ---- Pipeline::run(Some(&cache)) — pipeline.rs:340-372

extern "C++" {
namespace {

// Two independent 64-bit FNV-1a style streams: the keys never leave the process, so the hash only has to be
// collision-free in practice (the reference uses blake3 over bincode for the same purpose, hasher.rs:12-47).
struct ChainHash {
  uint64_t a = 0xcbf29ce484222325ull, b = 0x84222325cbf29ce4ull;
  void bytes(const void *p, size_t n) {
    const unsigned char *c = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) {
      a = (a ^ c[i]) * 0x100000001b3ull;
      b = (b ^ (c[i] + 0x9e)) * 0x9e3779b97f4a7c15ull;
      b ^= b >> 29;
    }
  }
  template <class T> void pod(const T &v) { bytes(&v, sizeof(v)); }
  void name(const char *s) { bytes(s, strlen(s)); }  // ImageOp::hash writes the op name first (pipeline.rs:88-92)
  ipb_cache::Key key() const { ipb_cache::Key k; k.a = a; k.b = b; return k; }
};

ipb_buffer *cache_get(ipb_cache *c, const ipb_cache::Key &k) {
  std::lock_guard<std::mutex> g(c->mu);
  for (size_t i = 0; i < c->lru.size(); i++)
    if (c->lru[i].key == k) {
      ipb_cache::Entry e = c->lru[i];
      c->lru.erase(c->lru.begin() + i);
      c->lru.push_back(e);  // most recently used
      c->hits++;
      ipb_buffer_retain(e.buf);
      return e.buf;
    }
  c->misses++;
  return nullptr;
}

void cache_put(ipb_cache *c, const ipb_cache::Key &k, ipb_buffer *buf) {
  const size_t bytes = buf->width * buf->height * buf->colors * 4;  // pipeline.rs:369
  std::lock_guard<std::mutex> g(c->mu);
  for (size_t i = 0; i < c->lru.size(); i++)
    if (c->lru[i].key == k) {
      ipb_buffer_release(c->lru[i].buf);
      c->bytes -= c->lru[i].bytes;
      c->lru.erase(c->lru.begin() + i);
      break;
    }
  if (bytes > c->max_bytes) return;  // would evict everything and still not fit
  while (c->bytes + bytes > c->max_bytes && !c->lru.empty()) {
    ipb_buffer_release(c->lru.front().buf);
    c->bytes -= c->lru.front().bytes;
    c->lru.erase(c->lru.begin());
  }
  ipb_buffer_retain(buf);
  ipb_cache::Entry e;
  e.key = k; e.buf = buf; e.bytes = bytes;
  c->lru.push_back(e);
  c->bytes += bytes;
}

}  // namespace
}  // extern "C++"

int ipb_cache_create(ipb_ctx *ctx, size_t max_bytes, ipb_cache **out) {
  if (!ctx || !out) return IPB_ERR_INVALID;
  ipb_cache *c = new (std::nothrow) ipb_cache();
  if (!c) return fail(ctx, IPB_ERR_NOMEM, "out of host memory");
  c->ctx = ctx;
  c->max_bytes = max_bytes;
  *out = c;
  return IPB_OK;
}

void ipb_cache_clear(ipb_cache *c) {
  if (!c) return;
  std::lock_guard<std::mutex> g(c->mu);
  for (auto &e : c->lru) ipb_buffer_release(e.buf);
  c->lru.clear();
  c->bytes = 0;
}

void ipb_cache_destroy(ipb_cache *c) {
  if (!c) return;
  ipb_cache_clear(c);
  delete c;
}

size_t ipb_cache_bytes(const ipb_cache *c) { return c ? c->bytes : 0; }
size_t ipb_cache_entries(const ipb_cache *c) { return c ? c->lru.size() : 0; }

int ipb_pipeline_run_cached(ipb_pipeline *p, ipb_cache *cache, ipb_buffer **out) {
  if (!p || !out) return IPB_ERR_INVALID;
  if (!cache) return ipb_pipeline_run(p, out);
  ipb_ctx *ctx = p->ctx;
  IPB_TRY(enter(ctx));
  if (cache->ctx != ctx) return fail(ctx, IPB_ERR_INVALID, "the cache belongs to another context (device buffers are per context)");
  if (p->image.width < 10 || p->image.height < 10) return fail(ctx, IPB_ERR_INVALID, "source smaller than 10x10");
  if (p->has_stripe) return fail(ctx, IPB_ERR_UNSUPPORTED, "pipeline_run on a stripe source: use output_8bit_stripe");
  negotiate(p, nullptr, nullptr);
  // the hash chain: settings first (pipeline.rs:346), then every op cumulatively (:350-361).  The source's identity is
  // hashed too (the reference leaves that to the caller: one cache per image): its address, shape, and a counter that
  // ipb_pipeline_set_source bumps — a ring buffer refilled with the next frame is a different image at the same address.
  // (Pixels rewritten in place WITHOUT set_source are not seen: like the reference, call set_source or clear the cache.)
  ChainHash h;
  const ipb_settings &st = p->settings;
  h.pod(st.maxwidth); h.pod(st.maxheight); h.pod(st.demosaic_width); h.pod(st.demosaic_height);
  h.pod(st.linear); h.pod(st.use_fastpath);
  h.pod(p->image.kind); h.pod(p->image.width); h.pod(p->image.height); h.pod(p->image.cpp); h.pod(p->image.data); h.pod(p->source_gen);
  ipb_cache::Key keys[8];
  const ipb_ops &o = p->ops;
  h.name("gofloat");
  h.pod(o.gofloat.crop_top); h.pod(o.gofloat.crop_right); h.pod(o.gofloat.crop_bottom); h.pod(o.gofloat.crop_left);
  h.pod(o.gofloat.is_cfa); h.pod(o.gofloat.blacklevels); h.pod(o.gofloat.whitelevels);
  keys[0] = h.key();
  h.name("demosaic"); h.bytes(o.demosaic.cfa, strnlen(o.demosaic.cfa, sizeof(o.demosaic.cfa)));
  keys[1] = h.key();
  h.name("rotatecrop");
  h.pod(o.rotatecrop.crop_top); h.pod(o.rotatecrop.crop_right); h.pod(o.rotatecrop.crop_bottom);
  h.pod(o.rotatecrop.crop_left); h.pod(o.rotatecrop.rotation);
  keys[2] = h.key();
  h.name("to_lab");
  h.pod(o.tolab.cam_to_xyz); h.pod(o.tolab.cam_to_xyz_normalized); h.pod(o.tolab.xyz_to_cam); h.pod(o.tolab.wb_coeffs);
  keys[3] = h.key();
  h.name("basecurve"); h.pod(o.basecurve.exposure); h.pod(o.basecurve.npoints);
  for (size_t i = 0; i < o.basecurve.npoints && i < IPB_MAX_CURVE_POINTS; i++) h.pod(o.basecurve.points[i]);
  keys[4] = h.key();
  h.name("from_lab");
  keys[5] = h.key();
  h.name("gamma");
  keys[6] = h.key();
  h.name("transform"); h.pod(o.transform.rotation); h.pod(o.transform.fliph); h.pod(o.transform.flipv);
  keys[7] = h.key();
  // the latest op whose output is cached (pipeline.rs:355-360)
  ipb_buffer *cur = nullptr;
  int startpos = 0;
  for (int i = 7; i >= 0 && !cur; i--) {
    cur = cache_get(cache, keys[i]);
    if (cur) startpos = i + 1;
  }
  p->last_startpos = startpos;
  p->last_ops_run = 8 - startpos;
  // the remaining ops, one kernel each, every result into the cache (pipeline.rs:364-371)
  for (int i = startpos; i < 8; i++) {
    ipb_buffer *next = nullptr;
    int rc;
    switch (i) {
      case 0: rc = ipb_gofloat_run(ctx, &p->ops.gofloat, &p->image, &next); break;
      case 1: rc = ipb_demosaic_run(ctx, &p->ops.demosaic, &p->settings, cur, &next); break;
      case 2: rc = ipb_rotatecrop_run(ctx, &p->ops.rotatecrop, cur, &next); break;
      case 3: rc = ipb_tolab_run(ctx, &p->ops.tolab, cur, &next); break;
      case 4: rc = ipb_basecurve_run(ctx, &p->ops.basecurve, cur, &next); break;
      case 5: rc = ipb_fromlab_run(ctx, cur, &next); break;
      case 6: rc = ipb_gamma_run(ctx, &p->settings, cur, &next); break;
      default: rc = ipb_transform_run(ctx, &p->ops.transform, cur, &next); break;
    }
    if (cur) ipb_buffer_release(cur);
    if (rc != IPB_OK) return rc;
    cur = next;
    cache_put(cache, keys[i], cur);
  }
  *out = cur;
  return IPB_OK;
}

void ipb_pipeline_last_run_info(const ipb_pipeline *p, int *startpos, int *ops_run) {
  if (!p) return;
  if (startpos) *startpos = p->last_startpos;
  if (ops_run) *ops_run = p->last_ops_run;
}

static bool ops_are_default_other(const ipb_pipeline *p) {  // Pipeline::default_ops (pipeline.rs:286-288)
  ipb_ops d;
  ipb_ops_default(&d, &p->image);
  const ipb_ops &a = p->ops;
  if (memcmp(&a.gofloat, &d.gofloat, sizeof(a.gofloat))) return false;
  if (strncmp(a.demosaic.cfa, d.demosaic.cfa, sizeof(a.demosaic.cfa))) return false;
  if (a.rotatecrop.crop_top != 0 || a.rotatecrop.crop_right != 0 || a.rotatecrop.crop_bottom != 0 ||
      a.rotatecrop.crop_left != 0 || a.rotatecrop.rotation != 0) return false;
  if (memcmp(&a.tolab, &d.tolab, sizeof(a.tolab))) return false;
  if (a.basecurve.exposure != 0 || a.basecurve.npoints != 0) return false;
  if (a.transform.rotation != 0 || a.transform.fliph || a.transform.flipv) return false;
  return true;
}

extern "C++" {
template <typename T>
static int output_impl(ipb_pipeline *p, ipb_cache *cache, T *dst, size_t cap, int dst_on_device, size_t *width, size_t *height) {
  if (!p || !dst) return IPB_ERR_INVALID;
  ipb_ctx *ctx = p->ctx;
  IPB_TRY(enter(ctx));
  if (p->image.width < 10 || p->image.height < 10) return fail(ctx, IPB_ERR_INVALID, "source smaller than 10x10");
  if (p->has_stripe) return fail(ctx, IPB_ERR_UNSUPPORTED, "stripe source: use output_8bit_stripe");
  const bool other = p->image.kind == IPB_SRC_RGB8 || p->image.kind == IPB_SRC_RGB16;
  const bool want8 = sizeof(T) == 1;

  if (other && p->settings.use_fastpath && ops_are_default_other(p)) {  // pipeline.rs:381-402 / :428-449
    const size_t w = p->image.width, h = p->image.height, n = w * h * 3;
    size_t nw, nh;
    ipb_scaling_size(w, h, p->settings.maxwidth, p->settings.maxheight, &nw, &nh);
    if (nw * nh * 3 > cap) return fail(ctx, IPB_ERR_INVALID, "destination too small: %zu < %zu", cap, nw * nh * 3);
    const bool same_depth = (p->image.kind == IPB_SRC_RGB8) == want8;
    const bool scale = nw != w || nh != h;
    DevSrc src;
    IPB_TRY(device_source(ctx, &p->image, &src));
    T *conv = nullptr;  // source raster at the output bit depth, on the device
    const T *raster = (const T *)src.ptr;
    if (!same_depth) {
      IPB_CUDA(ctx, ipb_malloc_async(ctx, (void **)&conv, (n ? n : 1) * sizeof(T)));
      if (want8) IPB_LAUNCH(ctx, launch_rgb16_to_8(ctx->stream, (const uint16_t *)src.ptr, n, (uint8_t *)conv));
      else IPB_LAUNCH(ctx, launch_rgb8_to_16(ctx->stream, (const uint8_t *)src.ptr, n, (uint16_t *)conv));
      raster = conv;
    }
    T *d = dst;
    if (!dst_on_device) IPB_CUDA(ctx, ipb_malloc_async(ctx, (void **)&d, nw * nh * 3 * sizeof(T)));
    if (scale) {
      XformGeom g;
      g.tl[0] = 0; g.tl[1] = 0; g.tr[0] = (long)w - 1; g.tr[1] = 0; g.bl[0] = 0; g.bl[1] = (long)h - 1;
      g.width = w; g.height = h; g.nwidth = nw; g.nheight = nh; g.components = 3;
      if (want8) IPB_LAUNCH(ctx, launch_transform_u8(ctx->stream, g, (const uint8_t *)raster, (uint8_t *)d));
      else IPB_LAUNCH(ctx, launch_transform_u16(ctx->stream, g, (const uint16_t *)raster, (uint16_t *)d));
    } else {
      IPB_CUDA(ctx, cudaMemcpyAsync(d, raster, n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (!dst_on_device) {
      IPB_CUDA(ctx, cudaMemcpyAsync(dst, d, nw * nh * 3 * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
      IPB_CUDA(ctx, cudaFreeAsync(d, ctx->stream));
    }
    if (conv) IPB_CUDA(ctx, cudaFreeAsync(conv, ctx->stream));
    release_source(ctx, &src);
    if (!dst_on_device) IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (width) *width = nw;
    if (height) *height = nh;
    return IPB_OK;
  }

  p->settings.linear = want8 ? 0 : 1;  // pipeline.rs:405 / :452
  if (cache) {  // self.run(cache) + the pack loop (pipeline.rs:406-414 / :453-461); the fast path above came first, as in the reference
    ipb_buffer *b;
    IPB_TRY(ipb_pipeline_run_cached(p, cache, &b));
    const size_t n = b->width * b->height * 3;
    int rc = n > cap ? fail(ctx, IPB_ERR_INVALID, "destination too small: %zu < %zu", cap, n) : pack_impl<T>(ctx, b, dst, dst_on_device);
    if (width) *width = b->width;
    if (height) *height = b->height;
    ipb_buffer_release(b);
    return rc;
  }
  size_t fw, fh;
  negotiate(p, &fw, &fh);
  FusedPlan plan;
  IPB_TRY(plan_fused(p, &plan));
  int flips[3];
  orientation_flips(&p->ops.transform, flips);
  const bool normal = !flips[0] && !flips[1] && !flips[2];
  if (plan.mode != kNotFused && normal) {
    const size_t n = plan.out_width * plan.out_height * 3;
    if (n > cap) return fail(ctx, IPB_ERR_INVALID, "destination too small: %zu < %zu", cap, n);
    IPB_TRY(run_fused_banded(p, plan, want8 ? kOutU8 : kOutU16, sizeof(T), 0, plan.out_height, dst, dst_on_device));
    if (!dst_on_device) IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (width) *width = plan.out_width;
    if (height) *height = plan.out_height;
    return IPB_OK;
  }
  ipb_buffer *b;
  IPB_TRY(ipb_pipeline_run(p, &b));
  const size_t n = b->width * b->height * 3;
  int rc = IPB_OK;
  if (n > cap) rc = fail(ctx, IPB_ERR_INVALID, "destination too small: %zu < %zu", cap, n);
  else rc = pack_impl<T>(ctx, b, dst, dst_on_device);
  if (width) *width = b->width;
  if (height) *height = b->height;
  ipb_buffer_release(b);
  return rc;
}

}  // extern "C++"
int ipb_pipeline_output_8bit(ipb_pipeline *p, uint8_t *dst, size_t dst_capacity, int dst_on_device, size_t *width,
                             size_t *height) {
  return output_impl<uint8_t>(p, nullptr, dst, dst_capacity, dst_on_device, width, height);
}
int ipb_pipeline_output_16bit(ipb_pipeline *p, uint16_t *dst, size_t dst_capacity, int dst_on_device, size_t *width,
                              size_t *height) {
  return output_impl<uint16_t>(p, nullptr, dst, dst_capacity, dst_on_device, width, height);
}
int ipb_pipeline_output_8bit_cached(ipb_pipeline *p, ipb_cache *cache, uint8_t *dst, size_t dst_capacity, int dst_on_device,
                                    size_t *width, size_t *height) {
  return output_impl<uint8_t>(p, cache, dst, dst_capacity, dst_on_device, width, height);
}
int ipb_pipeline_output_16bit_cached(ipb_pipeline *p, ipb_cache *cache, uint16_t *dst, size_t dst_capacity, int dst_on_device,
                                     size_t *width, size_t *height) {
  return output_impl<uint16_t>(p, cache, dst, dst_capacity, dst_on_device, width, height);
}

// ---- row stripes

static int stripe_plan(ipb_pipeline *p, FusedPlan *plan) {
  ipb_ctx *ctx = p->ctx;
  if (p->image.width < 10 || p->image.height < 10) return fail(ctx, IPB_ERR_INVALID, "source smaller than 10x10");
  // the 8-bit output's plan (settings.linear = false, pipeline.rs:405) without leaving a trace in the settings: this is
  // also reached from the pure size query ipb_pipeline_stripe_rows
  const int keep_linear = p->settings.linear;
  p->settings.linear = 0;
  negotiate(p, nullptr, nullptr);
  const int rc_plan = plan_fused(p, plan);
  p->settings.linear = keep_linear;
  IPB_TRY(rc_plan);
  int flips[3];
  orientation_flips(&p->ops.transform, flips);
  if (plan->mode == kNotFused || flips[0] || flips[1] || flips[2])
    return fail(ctx, IPB_ERR_UNSUPPORTED, "row stripes need the fused CFA path with a Normal orientation");
  return IPB_OK;
}

int ipb_pipeline_stripe_rows(ipb_pipeline *p, size_t out_row0, size_t out_row1, size_t *src_row0, size_t *src_row1) {
  if (!p || !src_row0 || !src_row1) return IPB_ERR_INVALID;
  FusedPlan plan;
  IPB_TRY(stripe_plan(p, &plan));
  if (out_row0 >= out_row1 || out_row1 > plan.out_height)
    return fail(p->ctx, IPB_ERR_INVALID, "output rows [%zu,%zu) outside the %zu-row result", out_row0, out_row1, plan.out_height);
  fused_src_rows(p, plan, out_row0, out_row1, src_row0, src_row1);
  return IPB_OK;
}

int ipb_stripe_plan(const ipb_ops *ops, const ipb_settings *settings, size_t width, size_t height, size_t out_row0,
                    size_t out_row1, size_t *src_row0, size_t *src_row1, size_t *out_width, size_t *out_height) {
  if (!ops) return IPB_ERR_INVALID;
  ipb_pipeline tmp;  // never touches a device: ctx stays null (failures land in the thread's create-error string)
  tmp.image.kind = IPB_SRC_RAW_U16;
  tmp.image.width = width;
  tmp.image.height = height;
  tmp.image.cpp = 1;
  tmp.ops = *ops;
  if (settings) tmp.settings = *settings;
  else tmp.settings.use_fastpath = 1;
  FusedPlan plan;
  IPB_TRY(stripe_plan(&tmp, &plan));
  if (out_width) *out_width = plan.out_width;
  if (out_height) *out_height = plan.out_height;
  if (out_row0 < out_row1) {
    if (out_row1 > plan.out_height || !src_row0 || !src_row1)
      return fail(nullptr, IPB_ERR_INVALID, "output rows [%zu,%zu) outside the %zu-row result", out_row0, out_row1, plan.out_height);
    fused_src_rows(&tmp, plan, out_row0, out_row1, src_row0, src_row1);
  }
  return IPB_OK;
}

int ipb_pipeline_set_stripe_source(ipb_pipeline *p, const ipb_source *rows, const ipb_stripe *stripe) {
  if (!p || !rows || !stripe) return IPB_ERR_INVALID;
  if (rows->kind != IPB_SRC_RAW_U16 || rows->cpp != 1 || rows->width != p->image.width)
    return fail(p->ctx, IPB_ERR_INVALID, "stripe rows must be u16 CFA rows of the pipeline's sensor width");
  if (stripe->full_height != p->image.height || stripe->src_row0 + rows->height > p->image.height)
    return fail(p->ctx, IPB_ERR_INVALID, "stripe does not fit the %zu-row frame", p->image.height);
  p->stripe_rows = *rows;
  p->stripe = *stripe;
  p->has_stripe = true;
  return IPB_OK;
}

int ipb_pipeline_output_8bit_stripe(ipb_pipeline *p, uint8_t *dst, size_t dst_capacity, int dst_on_device,
                                    size_t *width, size_t *rows) {
  if (!p || !dst) return IPB_ERR_INVALID;
  ipb_ctx *ctx = p->ctx;
  IPB_TRY(enter(ctx));
  if (!p->has_stripe) return fail(ctx, IPB_ERR_INVALID, "no stripe source set");
  FusedPlan plan;
  IPB_TRY(stripe_plan(p, &plan));
  const size_t r0 = p->stripe.out_row0, r1 = p->stripe.out_row1;
  if (r0 >= r1 || r1 > plan.out_height) return fail(ctx, IPB_ERR_INVALID, "stripe output rows [%zu,%zu) outside the %zu-row result", r0, r1, plan.out_height);
  const size_t n = (r1 - r0) * plan.out_width * 3;
  if (n > dst_capacity) return fail(ctx, IPB_ERR_INVALID, "destination too small: %zu < %zu", dst_capacity, n);
  IPB_TRY(run_fused_banded(p, plan, kOutU8, 1, r0, r1, dst, dst_on_device));
  if (!dst_on_device) IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (width) *width = plan.out_width;
  if (rows) *rows = r1 - r0;
  return IPB_OK;
}

// A batch of frames of identical geometry and parameters, device resident, frame k's source rows src_stride_rows * k rows
// after the pipeline's source (whole image or stripe rows) and its result dst_stride_bytes * k bytes after dst: the same
// bytes as nframes calls of output_8bit / output_8bit_stripe on shifted pointers.  One launch when the speculative
// full-resolution kernel applies (its tables, start-up and tail are paid once per batch), else one launch per frame.
int ipb_pipeline_output_8bit_batch(ipb_pipeline *p, size_t nframes, size_t src_stride_rows, uint8_t *dst, size_t dst_stride_bytes,
                                   size_t dst_capacity, size_t *width, size_t *rows) {
  if (!p || !dst) return IPB_ERR_INVALID;
  ipb_ctx *ctx = p->ctx;
  IPB_TRY(enter(ctx));
  ipb_source &src = p->has_stripe ? p->stripe_rows : p->image;
  if (!src.on_device) return fail(ctx, IPB_ERR_UNSUPPORTED, "output_8bit_batch: device-resident sources only");
  if (src_stride_rows < src.height) return fail(ctx, IPB_ERR_INVALID, "output_8bit_batch: frames overlap (%zu < %zu rows)", src_stride_rows, src.height);
  FusedPlan plan;
  size_t r0, r1;
  if (p->has_stripe) {
    IPB_TRY(stripe_plan(p, &plan));
    r0 = p->stripe.out_row0; r1 = p->stripe.out_row1;
    if (r0 >= r1 || r1 > plan.out_height) return fail(ctx, IPB_ERR_INVALID, "stripe output rows [%zu,%zu) outside the %zu-row result", r0, r1, plan.out_height);
  } else {
    if (p->image.width < 10 || p->image.height < 10) return fail(ctx, IPB_ERR_INVALID, "source smaller than 10x10");
    p->settings.linear = 0;  // pipeline.rs:405
    negotiate(p, nullptr, nullptr);
    IPB_TRY(plan_fused(p, &plan));
    int flips[3];
    orientation_flips(&p->ops.transform, flips);
    if (plan.mode == kNotFused || flips[0] || flips[1] || flips[2])
      return fail(ctx, IPB_ERR_UNSUPPORTED, "output_8bit_batch: only the fused raw CFA path with Normal orientation (use output_8bit per frame)");
    r0 = 0; r1 = plan.out_height;
  }
  const size_t n = (r1 - r0) * plan.out_width * 3;
  if (nframes > 0 && (dst_stride_bytes < n || (nframes - 1) * dst_stride_bytes + n > dst_capacity))
    return fail(ctx, IPB_ERR_INVALID, "output_8bit_batch: destination too small or strides overlap");
  if (width) *width = plan.out_width;
  if (rows) *rows = r1 - r0;
  if (nframes == 0) return IPB_OK;
  if (nframes > 1) {
    p->batch_n = nframes; p->batch_src_rows = src_stride_rows; p->batch_out_bytes = dst_stride_bytes; p->batch_done = false;
    const int rc = run_fused(p, plan, kOutU8, r0, r1, dst);
    const bool done = p->batch_done;
    p->batch_n = 0; p->batch_done = false;
    if (rc == IPB_OK && done) return IPB_OK;
    if (rc != IPB_OK && rc != IPB_ERR_UNSUPPORTED) return rc;
  }
  // frame by frame on shifted pointers
  const void *base = src.data;
  int rc = IPB_OK;
  for (size_t k = 0; k < nframes && rc == IPB_OK; k++) {
    src.data = (const uint8_t *)base + k * src_stride_rows * src.width * sizeof(uint16_t);
    rc = run_fused(p, plan, kOutU8, r0, r1, dst + k * dst_stride_bytes);
  }
  src.data = base;
  return rc;
}

int ipb_pipeline_set_band_mb(ipb_pipeline *p, int megabytes) {
  if (!p || megabytes < 0) return IPB_ERR_INVALID;
  p->band_mb = megabytes;
  return IPB_OK;
}

int ipb_pipeline_set_speculative(ipb_pipeline *p, int on) {
  if (!p) return IPB_ERR_INVALID;
  p->spec = on ? 1 : 0;
  return IPB_OK;
}

int ipb_ctx_set_spec(ipb_ctx *ctx, float delta, int threads) {
  IPB_TRY(enter(ctx));
  if (!(delta >= 0.0f) || (threads != 512 && threads != 1024)) return fail(ctx, IPB_ERR_INVALID, "set_spec: delta >= 0, threads 512 or 1024");
  ctx->spec_delta_override = delta;
  ctx->spec_threads = threads;
  return IPB_OK;
}

int ipb_ctx_spec_stats(ipb_ctx *ctx, unsigned long long out[4], int reset) {
  IPB_TRY(enter(ctx));
  if (!out) return IPB_ERR_INVALID;
  unsigned long long h[8] = {0};
  IPB_CUDA(ctx, cudaMemcpyAsync(h, ctx->spec_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  if (reset) IPB_CUDA(ctx, cudaMemsetAsync(ctx->spec_stats, 0, sizeof(h), ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint32_t db = 0, mb = 0;
  const float d = ctx->spec_tab_state > 0 ? ctx->spec_tab.delta : 0.0f;
  memcpy(&db, &d, 4);
  memcpy(&mb, &ctx->mufu_cbrt_err, 4);
  out[0] = h[0]; out[1] = h[4]; out[2] = db; out[3] = mb;
  return IPB_OK;
}

int ipb_spec_tables(const ipb_ops *ops, float mufu_rel_err, float delta_override, uint32_t *g8a, float *thresholds, float one[3],
                    uint32_t wmul[3], uint32_t *amb_t, float delta[4]) {
  if (!ops || !g8a || !thresholds || !one || !wmul || !amb_t || !delta) return IPB_ERR_INVALID;
  static std::once_flag once;
  std::call_once(once, [] { if (!g_gamma8_ok && g_gamma8.empty()) g_gamma8_ok = build_gamma8(tables().fwd, &g_gamma8); });
  if (!g_gamma8_ok) return IPB_ERR_UNSUPPORTED;
  ColorParams P;
  memset(&P, 0, sizeof(P));
  fill_tolab(&P, &ops->tolab, 0);
  P.use_e = 0;
  if (!build_spline(&ops->basecurve, &P.sp)) return IPB_ERR_INVALID;
  std::vector<float> thr;
  for (const Gamma8Entry &g : g_gamma8)
    if (g.thr <= 1.0f) thr.push_back(g.thr);
  std::vector<uint32_t> tab;
  std::vector<float2> stab;
  SpecParams c;
  const float black = ops->gofloat.blacklevels[0], range = ops->gofloat.whitelevels[0] - black;
  if (!spec_build(P, black, range, mufu_rel_err, delta_override, thr, &tab, &stab, &c, delta)) return IPB_ERR_UNSUPPORTED;
  if (thr.size() != 255 || tab.size() != (size_t)kSpecG8Entries) return IPB_ERR_UNSUPPORTED;
  memcpy(g8a, tab.data(), tab.size() * sizeof(uint32_t));
  memcpy(thresholds, thr.data(), 255 * sizeof(float));
  for (int k = 0; k < 3; k++) { one[k] = c.one[k]; wmul[k] = c.wmul[k]; }
  *amb_t = c.amb_t;
  return IPB_OK;
}

int ipb_scaled_division_check(size_t width, size_t height, size_t nwidth, size_t nheight) {
  return scaled_skip_division_exact(width, height, nwidth, nheight) ? 1 : 0;
}

int ipb_spec_bound(const ipb_ops *ops, float mufu_rel_err, float *delta) {
  if (!ops || !delta) return IPB_ERR_INVALID;
  static std::once_flag once;
  std::call_once(once, [] { if (!g_gamma8_ok && g_gamma8.empty()) g_gamma8_ok = build_gamma8(tables().fwd, &g_gamma8); });
  if (!g_gamma8_ok) return IPB_ERR_UNSUPPORTED;
  ColorParams P;
  memset(&P, 0, sizeof(P));
  fill_tolab(&P, &ops->tolab, 0);
  P.use_e = 0;
  if (!build_spline(&ops->basecurve, &P.sp)) return IPB_ERR_INVALID;
  std::vector<float> thr;
  for (const Gamma8Entry &g : g_gamma8)
    if (g.thr <= 1.0f) thr.push_back(g.thr);
  std::vector<uint32_t> g8a;
  std::vector<float2> stab;
  SpecParams c;
  const float black = ops->gofloat.blacklevels[0], range = ops->gofloat.whitelevels[0] - black;
  float dl[4];
  if (!spec_build(P, black, range, mufu_rel_err, 0.0f, thr, &g8a, &stab, &c, dl)) return IPB_ERR_UNSUPPORTED;
  for (int k = 0; k < 4; k++) delta[k] = dl[k];
  return IPB_OK;
}

int ipb_pipeline_spec_probe(ipb_pipeline *p, float *max_dev, double *mean_dev, float *delta) {
  if (!p || !max_dev || !mean_dev || !delta) return IPB_ERR_INVALID;
  ipb_ctx *ctx = p->ctx;
  IPB_TRY(enter(ctx));
  if (p->has_stripe) return fail(ctx, IPB_ERR_UNSUPPORTED, "spec_probe: whole-frame sources only");
  const int keep_linear = p->settings.linear;
  p->settings.linear = 0;
  negotiate(p, nullptr, nullptr);
  FusedPlan plan;
  int rc = plan_fused(p, &plan);
  ColorParams P;
  if (rc == IPB_OK) rc = plan.mode == kFusedFull ? fill_color_params(p, plan, &P) : fail(ctx, IPB_ERR_UNSUPPORTED, "spec_probe: not the full-resolution fused path");
  p->settings.linear = keep_linear;
  IPB_TRY(rc);
  const uint16_t *raw = (const uint16_t *)p->image.data;
  const size_t bytes = p->image.width * p->image.height * sizeof(uint16_t);
  if (!p->image.on_device) {
    IPB_TRY(ensure_stage(ctx, &p->stage_in, &p->stage_in_bytes, bytes));
    IPB_CUDA(ctx, cudaMemcpyAsync(p->stage_in, raw, bytes, cudaMemcpyHostToDevice, ctx->stream));
    raw = (const uint16_t *)p->stage_in;
  }
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.raw = raw; a.raw_pitch = p->image.width; a.src_row0 = 0; a.src_rows = p->image.height;
  a.crop_x = plan.crop_x; a.crop_y = plan.crop_y; a.width = plan.width; a.height = plan.height;
  a.out_row0 = 0; a.out_row1 = plan.height; a.out_width = plan.width; a.out_height = plan.height;
  a.out_kind = kOutU8;
  a.black = p->ops.gofloat.blacklevels[0];
  a.range = p->ops.gofloat.whitelevels[0] - a.black;
  a.range_rc = 1.0f / a.range;
  a.exact_rc = golevel_rc_exact(a.black, a.range, a.range_rc) ? 1 : 0;
  a.lut_lab = ctx->lut_lab; a.lut_gamma = ctx->lut_gamma; a.lut_gamma8 = ctx->lut_gamma8;
  a.cbrt_tab = ctx->cbrt_tab;
  a.use_tma = 1;
  bool use = false;
  if (a.exact_rc && spec_supported(a, plan.cfa, P)) IPB_TRY(ensure_spec_tables(ctx, P, a.black, a.range, &use));
  if (!use) return fail(ctx, IPB_ERR_UNSUPPORTED, "spec_probe: these parameters do not take the speculative path");
  IPB_CUDA(ctx, cudaMemsetAsync(ctx->spec_stats + 1, 0, 3 * sizeof(unsigned long long), ctx->stream));
  IPB_LAUNCH(ctx, launch_spec_probe(ctx->stream, a, plan.cfa, P, ctx->spec_tab, ctx->sm_count));
  unsigned long long h[4] = {0};
  IPB_CUDA(ctx, cudaMemcpyAsync(h, ctx->spec_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint32_t mb = (uint32_t)h[1];
  memcpy(max_dev, &mb, 4);
  *mean_dev = h[2] ? (double)h[3] / 1099511627776.0 / (double)h[2] : 0.0;
  *delta = ctx->spec_tab.delta;
  return IPB_OK;
}

int ipb_pipeline_set_tma(ipb_pipeline *p, int use_tma) {
  if (!p) return IPB_ERR_INVALID;
  p->use_tma = use_tma != 0;
  return IPB_OK;
}

int ipb_selftest_gamma8(ipb_ctx *ctx, unsigned long long *mismatches) {
  IPB_TRY(enter(ctx));
  if (!mismatches) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (!ctx->lut_gamma8) return fail(ctx, IPB_ERR_UNSUPPORTED, "the 8-bit gamma threshold table failed its build-time checks");
  unsigned long long *d;
  IPB_CUDA(ctx, ipb_malloc_async(ctx, (void **)&d, sizeof(*d)));
  IPB_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(*d), ctx->stream));
  IPB_LAUNCH(ctx, launch_gamma8_selftest(ctx->stream, ctx->lut_gamma, ctx->lut_gamma8, d));
  IPB_CUDA(ctx, cudaMemcpyAsync(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost, ctx->stream));
  IPB_CUDA(ctx, cudaFreeAsync(d, ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IPB_OK;
}

int ipb_gamma_pack_8bit(ipb_ctx *ctx, const float *in, size_t n, uint8_t *out) {
  IPB_TRY(enter(ctx));
  if (!in || !out) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  if (!ctx->lut_gamma8) return fail(ctx, IPB_ERR_UNSUPPORTED, "the 8-bit gamma threshold table failed its build-time checks");
  float *d;
  IPB_CUDA(ctx, ipb_malloc_async(ctx, (void **)&d, (n ? n : 1) * 5));
  uint8_t *o = (uint8_t *)(d + n);
  IPB_CUDA(ctx, cudaMemcpyAsync(d, in, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  IPB_LAUNCH(ctx, launch_gamma8_pack(ctx->stream, ctx->lut_gamma8, d, n, o));
  IPB_CUDA(ctx, cudaMemcpyAsync(out, o, n, cudaMemcpyDeviceToHost, ctx->stream));
  IPB_CUDA(ctx, cudaFreeAsync(d, ctx->stream));
  IPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return IPB_OK;
}

int ipb_synth_cfa_u16(ipb_ctx *ctx, uint64_t seed, size_t width, size_t row0, size_t rows, uint16_t *dptr) {
  IPB_TRY(enter(ctx));
  if (!dptr) return fail(ctx, IPB_ERR_INVALID, "null pointer");
  IPB_LAUNCH(ctx, launch_synth(ctx->stream, seed, width, row0, rows, dptr));
  return IPB_OK;
}

}  // extern "C"
