// ipb_spec.h — parameters and tables of the speculative 8-bit kernel (ipb_spec.cu), shared with the host layer that
// builds them (ipb_host.cu: build_spec_tables, spec_error_bound).  Not part of the public ABI.
#pragma once
#include <vector>

#include "ipb_internal.h"

namespace ipb {

constexpr int kSpecG8Entries = 8192;          // gamma table segments: F >> 10, F = round(v * kSpecFScale)
constexpr double kSpecFScale = 8388608.0 - 1024.0;  // 2^23 - 2^10: v = 1 lands on the first value of segment 8191
constexpr int kSpecSTabN = 2048;              // basecurve table: segments over fy in [16/116, 1]
constexpr int kSpecSTabEntries = kSpecSTabN + 4;  // + one guard below, two above, one of padding (16-byte multiple)
constexpr float kSpecYMin = -0.03f;           // pixels whose Y ratio lies below are outside the certified domain: recomputed
constexpr uint32_t kSpecSmemBase = 0x400;     // shared-window address of dynamic shared memory (probed at context creation)

struct SpecParams {
  // ---- geometry / source (filled per launch, as FullParams of ipb_fused.cu)
  const uint16_t *raw;
  long long raw_pitch;
  int src_row0, src_rows;
  int crop_x, crop_y;
  int width, height;
  int out_row0, out_row1;
  uint8_t *out;
  int tiles_x, tiles_y;
  int nframes;                // frames of identical geometry in this launch (a batch), >= 1
  int frame_src_rows;         // ... source rows from one frame to the next
  long long frame_out_bytes;  // ... output bytes from one frame to the next
  int pw, ph;                 // CFA period (generic-pattern variant)
  uint32_t rcp_pw, rcp_ph;    // floor(2^32 / period) + 1: n % d = n - mulhi(n, rcp) * d while n * d < 2^32
  // ---- exact path (fix-ups)
  float black, range, range_rc;
  int exact_rc;
  const float2 *lut_lab, *lut_gamma;
  const float *cbrt_tab;  // host cbrtf of every float in (1, 1.5] (ipb_host.cu ensure_cbrt_table)
  // ---- cheap path: tables
  const uint32_t *g8a;   // kSpecG8Entries words: (byte << 24) + (2^24 - thrF) + deltaF - 0x3F800000
  const float2 *stab;    // kSpecSTabEntries {intercept, slope}
  const float *thr8;     // the 255 thresholds of output8bit(apply_srgb_gamma(v)), ascending (exact path)
  unsigned long long *stats;  // optional: [0] pixels recomputed, [1] probe max |diff| bits, [2] probe count, [3] probe sum
  // ---- cheap path: constants (scalars: packed instructions take them as broadcast operands)
  float sub_a, sub_b;    // level mapping: -(2^23 + black), 0 for an integral black level, else -2^23, -black
  float lim_r, lim_b;    // clip limits 1/mul[0], 1/mul[2]
  float m[3][3];         // cam -> XYZ ratio: cm[i][j] * mul[j] / white[i]
  float ro[3][3];        // XYZ ratio -> linear sRGB: rgbm[i][j] * white[j]
  float s_scale, s_off;  // u = sat(fy * s_scale + s_off), key = floor(u * (N + 2))
  float y_min;           // kSpecYMin
  uint32_t bias58;       // 0x58000000 = (0x4B000000 << 3) mod 2^32, as a run-time value (keeps the table address one LEA)
  // per-channel certificate: u_c = v * (1 - 2^-13) + one[c] carries F + deltaF_c, so the low 24 bits of (entry + bits(u_c))
  // are r = F - thrF + deltaF_c, and the channel is within deltaF_c of a threshold iff r <= 2 * deltaF_c.  The three
  // tests share one comparison: r * wmul[c] <= amb_t with wmul[c] = 256 * floor(4 * deltaF_max / deltaF_c) (the factor
  // 256 shifts the byte field out of the 32-bit product) and amb_t = 256 * 4 * 2 * deltaF_max.
  float one[3];
  uint32_t wmul[3];
  uint32_t amb_t;
  int dbg;               // timing experiments only (IPB_SPEC_DBG): 1 = skip the recomputation, 2 = skip the queueing too
};

struct SpecTables {
  SpecParams consts;     // the constant part of SpecParams
  const uint32_t *g8a;
  const float2 *stab;
  const float *thr8;
  unsigned long long *stats;
  float delta;           // the certified bound on |cheap - exact| in linear units (largest channel) of this table set
  float delta_ch[3];     // per output channel
};

// host: tables, folded constants and the certified bound for one parameter set (ipb_spec_host.cu).  `thresholds` are the
// 255 values of v where output8bit(apply_srgb_gamma(v)) steps up (ipb_host.cu build_gamma8).  false: these parameters
// cannot take the speculative path (the caller launches k_fused_full instead).
bool spec_build(const ColorParams &P, float black, float range, float mufu_rel_err, float delta_override,
                const std::vector<float> &thresholds, std::vector<uint32_t> *g8a, std::vector<float2> *stab,
                SpecParams *consts, float delta_out[4]);
bool spec_supported(const FusedArgs &a, const CfaDev &cfa, const ColorParams &P);
cudaError_t launch_fused_spec8(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                               const SpecTables &T, int sm_count, int threads);
// the same for a down-scaled RGB Bayer frame (scaled_demosaic): k_spec8_scaled
bool spec_scaled_supported(const FusedArgs &a, const CfaDev &cfa, const ColorParams &P);
cudaError_t launch_scaled_spec8(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                                const SpecTables &T, int sm_count);
cudaError_t launch_spec_probe(cudaStream_t s, const FusedArgs &a, const CfaDev &cfa, const ColorParams &P,
                              const SpecTables &T, int sm_count);
cudaError_t launch_spec_selftest(cudaStream_t s, unsigned int *out2);
const char *spec_last_error();

}  // namespace ipb
