// ipb_fused.cu — the fused raw -> sRGB kernels (the roofline kernels of the hot path).
//
//   k_fused_full    gofloat (gofloat.rs:122-130) + demosaic::full (demosaic.rs:67-119) + to_lab + basecurve +
//                   from_lab + gamma (+ output8bit/output16bit pack) for a full-resolution CFA frame:
//                   u16 in (2 B/px), u8x3 / u16x3 / f32x3 out, nothing else touches HBM.
//   k_fused_scaled  the same chain with scaled_demosaic (scaling.rs:132-145) in place of full() — the branch
//                   OpDemosaic::run takes for scale >= 2 (Bayer) / 3 (X-Trans), demosaic.rs:47-50.
//
// Both are persistent kernels: one CTA per SM, both 8192-entry {v,dv} tables (128 KB) resident in shared
// memory for the whole launch, CTAs walking tiles round-robin.  Arithmetic is the per-pixel code of
// ipb_device.cuh, compiled -fmad=false: results are bit-identical to the reference's f32 arithmetic.
#include "ipb_internal.h"

namespace ipb {

namespace {

constexpr int kTW = 256;                  // tile width in pixels (64 four-pixel tasks per tile row = 2 warps)
constexpr int kTH = 16;                   // tile height
constexpr int kNT = 512;                  // threads per CTA
constexpr int kTileStride = kTW + 8;      // floats per tile row: frame col tx0-4 .. tx0+kTW+3 (col tx0 at index 4)
constexpr int kTileRows = kTH + 2;        // + one halo row above and below
constexpr int kMaxPatPos = 144;           // largest CFA period (12x12)

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
  const float2 *lut_lab, *lut_gamma;
  int tiles_x, tiles_y;
  int pw, ph;                   // CFA period
  int bayer;                    // 2x2 RGB Bayer: specialised interior path
};

struct Smem {
  float2 lut_lab[kLutEntries];
  float2 lut_gamma[kLutEntries];
  float tile[kTileRows * kTileStride];
  uint2 taps[kMaxPatPos];       // per pattern position: 9-bit tap masks of colours 0..3, 16 bits each
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void load_luts(float2 *s_lab, float2 *s_gam, const float2 *lab, const float2 *gam) {
  const uint4 *a = reinterpret_cast<const uint4 *>(lab);
  const uint4 *b = reinterpret_cast<const uint4 *>(gam);
  uint4 *da = reinterpret_cast<uint4 *>(s_lab);
  uint4 *db = reinterpret_cast<uint4 *>(s_gam);
  for (int i = threadIdx.x; i < kLutEntries / 2; i += blockDim.x) {
    da[i] = __ldg(a + i);
    db[i] = __ldg(b + i);
  }
}

// demosaic.rs:77-90 for every position of the CFA period: which of the nine 3x3 taps feed which colour bin.
// Taps of the centre's own colour other than the centre itself go to the discarded fifth bin.
__device__ __forceinline__ void build_taps(Smem &sm, const CfaDev &cfa, int pw, int ph) {
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

// ---------------------------------------------------------------------------------------- full resolution

template <int OUT>
__global__ void __launch_bounds__(kNT, 1)
k_fused_full(const __grid_constant__ FullParams p, const __grid_constant__ CfaDev cfa,
             const __grid_constant__ ColorParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;

  load_luts(sm.lut_lab, sm.lut_gamma, p.lut_lab, p.lut_gamma);
  build_taps(sm, cfa, p.pw, p.ph);
  const LutShared lab{smem_u32(sm.lut_lab)}, gam{smem_u32(sm.lut_gamma)};

  // Bayer phase (cropped-frame coordinates): colour of the pixel at (row&1, col&1)
  const int c00 = cfa.pat[0], c10 = cfa.pat[48];

  const int ntiles = p.tiles_x * p.tiles_y;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tyi = t / p.tiles_x, txi = t - tyi * p.tiles_x;
    const int ty0 = p.out_row0 + tyi * kTH, tx0 = txi * kTW;
    __syncthreads();  // previous tile fully consumed (and, first time round, tables complete)

    // ---- stage the tile: gofloat once per sensor pixel, zero outside the frame
    for (int i = tid; i < kTileRows * (kTW + 2); i += kNT) {
      const int r = i / (kTW + 2), c = i - r * (kTW + 2);
      const int y = ty0 - 1 + r, x = tx0 - 1 + c;
      float v = 0.0f;
      if (y >= 0 && y < p.height && x >= 0 && x < p.width) {
        const int sr = y + p.crop_y - p.src_row0;
        if (sr >= 0 && sr < p.src_rows) {
          const uint16_t rawv = __ldg(p.raw + (long long)sr * p.raw_pitch + p.crop_x + x);
          v = golevel((float)rawv, p.black, p.range, p.range_rc, p.exact_rc);
        }
      }
      sm.tile[r * kTileStride + c + 3] = v;
    }
    __syncthreads();

    // ---- four-pixel tasks: 64 per tile row, consecutive lanes on consecutive tasks
    for (int task = tid; task < (kTW / 4) * kTH; task += kNT) {
      const int r = task / (kTW / 4), q = task - r * (kTW / 4);
      const int y = ty0 + r, x0 = tx0 + 4 * q;
      const bool live = y < p.out_row1 && x0 < p.width;
      const int npx = live ? min(4, p.width - x0) : 0;
      const bool interior = y >= 1 && y <= p.height - 2 && x0 >= 1 && x0 + 4 <= p.width - 1;
      const bool fast = p.bayer && __all_sync(0xffffffffu, !live || interior);

      // 3 x 6 window: rows y-1..y+1, cols x0-1..x0+4
      float w[3][6];
      const float *tp = sm.tile + r * kTileStride + 4 * q + 3;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float4 mid = *reinterpret_cast<const float4 *>(tp + k * kTileStride + 1);
        w[k][0] = tp[k * kTileStride];
        w[k][1] = mid.x; w[k][2] = mid.y; w[k][3] = mid.z; w[k][4] = mid.w;
        w[k][5] = tp[k * kTileStride + 5];
      }

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
      } else {
        const int pr = y % p.ph;
        int pc = x0 % p.pw;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int x = x0 + j;
          uint32_t valid = 0x1ffu;
          if (y == 0) valid &= ~0x007u;
          if (y == p.height - 1) valid &= ~0x1c0u;
          if (x == 0) valid &= ~0x049u;
          if (x >= p.width - 1) valid &= ~0x124u;
          const uint2 mm = sm.taps[pr * p.pw + pc];
          pc = (pc + 1 == p.pw) ? 0 : pc + 1;
          float v[9];
#pragma unroll
          for (int k = 0; k < 3; k++) { v[k * 3] = w[k][j]; v[k * 3 + 1] = w[k][j + 1]; v[k * 3 + 2] = w[k][j + 2]; }
          cr[j] = bin_mean(mm.x & 0xffffu & valid, v);
          cg[j] = bin_mean((mm.x >> 16) & valid, v);
          cb[j] = bin_mean(mm.y & 0xffffu & valid, v);
          ce[j] = bin_mean((mm.y >> 16) & valid, v);
        }
      }

      float orr[4], og[4], ob[4];
#pragma unroll
      for (int j = 0; j < 4; j++) color_chain(P, lab, gam, cr[j], cg[j], cb[j], ce[j], orr[j], og[j], ob[j]);
      if (live) store_px4<OUT>(p.out, (size_t)(y - p.out_row0) * p.width + x0, npx, orr, og, ob);
    }
  }
}

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
};

struct SmemScaled {
  float2 lut_lab[kLutEntries];
  float2 lut_gamma[kLutEntries];
  uint8_t pat[48 * 48];
};

__device__ __forceinline__ int f2i_sat(float f) {  // Rust `f as usize` for the values met here (>= 0, < 2^31)
  return (int)min(__float2uint_rz(f), 0x7fffffffu);
}

template <int OUT>
__global__ void __launch_bounds__(kNT, 1)
k_fused_scaled(const __grid_constant__ ScaledParams p, const __grid_constant__ CfaDev cfa,
               const __grid_constant__ ColorParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemScaled &sm = *reinterpret_cast<SmemScaled *>(smem_raw);
  load_luts(sm.lut_lab, sm.lut_gamma, p.lut_lab, p.lut_gamma);
  for (int i = threadIdx.x; i < 48 * 48; i += blockDim.x) sm.pat[i] = cfa.pat[i];
  __syncthreads();
  const LutShared lab{smem_u32(sm.lut_lab)}, gam{smem_u32(sm.lut_gamma)};

  const long long npix = (long long)(p.out_row1 - p.out_row0) * p.nwidth;
  for (long long idx = (long long)blockIdx.x * kNT + threadIdx.x; idx < npix; idx += (long long)gridDim.x * kNT) {
    const int row = p.out_row0 + (int)(idx / p.nwidth), col = (int)(idx % p.nwidth);
    const float frow = (float)row, frow1 = (float)(row + 1), fcol = (float)col, fcol1 = (float)(col + 1);
    // scaling.rs:77-89 with topleft = (0,0), skip_x_y = skip_y_x = 0
    const float rfrom_x = 0.0f + 0.0f * frow;
    const float rto_x = 0.0f + 0.0f * frow1;
    const float rfrom_y = 0.0f + p.skip_y * frow;
    const float rto_y = 0.0f + p.skip_y * frow1;
    const float rcenter_x = 0.0f + (0.0f * frow) + __fdiv_rn(0.0f, 2.0f) - 0.5f;
    const float rcenter_y = 0.0f + (p.skip_y * frow) + __fdiv_rn(p.skip_y, 2.0f) - 0.5f;
    const int from_x = min(p.width - 1, f2i_sat(floorf(rfrom_x + (p.skip_x * fcol))));
    const int to_x = min(p.width - 1, f2i_sat(floorf(rto_x + (p.skip_x * fcol1))));
    const int from_y = min(p.height - 1, f2i_sat(floorf(rfrom_y + (0.0f * fcol))));
    const int to_y = min(p.height - 1, f2i_sat(floorf(rto_y + (0.0f * fcol1))));
    const float center_x = rcenter_x + (p.skip_x * fcol) + __fdiv_rn(p.skip_x, 2.0f);
    const float center_y = rcenter_y + (0.0f * fcol) + __fdiv_rn(0.0f, 2.0f);

    float sums[4] = {0.f, 0.f, 0.f, 0.f}, counts[4] = {0.f, 0.f, 0.f, 0.f};
    for (int y = from_y; y <= to_y; y++) {
      const float delta_y = __fdiv_rn((float)y - center_y, p.skip_y);
      const float dy2 = delta_y * delta_y;
      const uint16_t *rowp = p.raw + (long long)(y + p.crop_y - p.src_row0) * p.raw_pitch + p.crop_x;
      const uint8_t *prow = sm.pat + (y % 48) * 48;
      for (int x = from_x; x <= to_x; x++) {
        const float delta_x = __fdiv_rn((float)x - center_x, p.skip_x);
        float factor = 1.0f - (delta_x * delta_x) - dy2;
        factor = factor < 0.0f ? 0.0f : factor;
        const int c = prow[x % 48];
        const float v = golevel((float)__ldg(rowp + x), p.black, p.range, p.range_rc, p.exact_rc) * factor;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (c == k) { sums[k] += v; counts[k] += factor; }
      }
    }
    float px[4];
#pragma unroll
    for (int k = 0; k < 4; k++) px[k] = counts[k] > 0.0f ? __fdiv_rn(sums[k], counts[k]) : 0.0f;
    float r[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    color_chain(P, lab, gam, px[0], px[1], px[2], px[3], r[0], g[0], b[0]);
    store_px4<OUT>(p.out, (size_t)(idx), 1, r, g, b);
  }
}

thread_local const char *g_fused_err = "";

template <class K>
cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

const char *fused_last_error() { return g_fused_err; }

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
  p.lut_lab = a.lut_lab; p.lut_gamma = a.lut_gamma;
  p.tiles_x = (p.width + kTW - 1) / kTW;
  p.tiles_y = (p.out_row1 - p.out_row0 + kTH - 1) / kTH;
  p.pw = cfa.width; p.ph = cfa.height;
  p.bayer = is_rgb_bayer(cfa) ? 1 : 0;
  const int ntiles = p.tiles_x * p.tiles_y;
  const int grid = ntiles < sm_count ? ntiles : sm_count;
  const size_t smem = sizeof(Smem);
  cudaError_t e;
  switch (a.out_kind) {
    case kOutF32:
      if ((e = set_smem(k_fused_full<kOutF32>, smem)) != cudaSuccess) return e;
      k_fused_full<kOutF32><<<grid, kNT, smem, s>>>(p, cfa, P);
      break;
    case kOutU8:
      if ((e = set_smem(k_fused_full<kOutU8>, smem)) != cudaSuccess) return e;
      k_fused_full<kOutU8><<<grid, kNT, smem, s>>>(p, cfa, P);
      break;
    default:
      if ((e = set_smem(k_fused_full<kOutU16>, smem)) != cudaSuccess) return e;
      k_fused_full<kOutU16><<<grid, kNT, smem, s>>>(p, cfa, P);
      break;
  }
  return cudaGetLastError();
}

cudaError_t launch_fused_scaled(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                                int sm_count) {
  if (a.out_row1 <= a.out_row0 || a.out_width == 0) return cudaSuccess;
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
  const long long npix = (long long)(p.out_row1 - p.out_row0) * p.nwidth;
  long long blocks = (npix + kNT - 1) / kNT;
  const int grid = (int)(blocks < sm_count ? blocks : sm_count);
  const size_t smem = sizeof(SmemScaled);
  cudaError_t e;
  switch (a.out_kind) {
    case kOutF32:
      if ((e = set_smem(k_fused_scaled<kOutF32>, smem)) != cudaSuccess) return e;
      k_fused_scaled<kOutF32><<<grid, kNT, smem, s>>>(p, cfa, P);
      break;
    case kOutU8:
      if ((e = set_smem(k_fused_scaled<kOutU8>, smem)) != cudaSuccess) return e;
      k_fused_scaled<kOutU8><<<grid, kNT, smem, s>>>(p, cfa, P);
      break;
    default:
      if ((e = set_smem(k_fused_scaled<kOutU16>, smem)) != cudaSuccess) return e;
      k_fused_scaled<kOutU16><<<grid, kNT, smem, s>>>(p, cfa, P);
      break;
  }
  return cudaGetLastError();
}

}  // namespace ipb
