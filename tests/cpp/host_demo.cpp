// host_demo.cpp — a compiled (C++17) caller of the C ABI through include/imagepipe_b200.hpp, the mirror of the
// reference's Pipeline / ImageOp surface.  Built and run by tests/test_cpp_host.py.
//   host_demo sizes                                   host-only checks (no GPU): size negotiation, defaults
//   host_demo run W H SEED OUT.bin [MAXWIDTH]         raw RGGB frame -> Pipeline::output_8bit, written to OUT.bin,
//                                                     and the same frame op by op through ImageOp::run into OUT.bin.ops
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "imagepipe_b200.hpp"

using namespace imagepipe;

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// the literal camera metadata of tests/common.py
static void fill_ops(PipelineOps ops) {
  for (int i = 0; i < 4; i++) { ops.gofloat.blacklevels[i] = 512.0f; ops.gofloat.whitelevels[i] = 16383.0f; }
  ops.gofloat.is_cfa = 1;
  std::strcpy(ops.demosaic.cfa, "RGGB");
  const float m[3][4] = {{0.6097f, 0.2053f, 0.1355f, 0.0f}, {0.2762f, 0.8149f, -0.0911f, 0.0f}, {0.0297f, -0.1206f, 1.1797f, 0.0f}};
  std::memcpy(ops.tolab.cam_to_xyz, m, sizeof(m));
  std::memcpy(ops.tolab.cam_to_xyz_normalized, m, sizeof(m));
  ops.tolab.wb_coeffs[0] = 2.0f; ops.tolab.wb_coeffs[1] = 1.0f; ops.tolab.wb_coeffs[2] = 1.5f; ops.tolab.wb_coeffs[3] = NAN;
  ops.basecurve.exposure = 0.0f; ops.basecurve.npoints = 1;
  ops.basecurve.points[0][0] = 0.5f; ops.basecurve.points[0][1] = 0.6f;
}

static int sizes() {
  size_t w, h;
  ipb_scaling_size(6000, 4000, 1500, 1000, &w, &h);
  if (w != 1500 || h != 1000) return 1;
  ipb_scaling_size(128, 64, 256, 0, &w, &h);  // never upscales (scaling.rs:15-16)
  if (w != 128 || h != 64) return 2;
  OpGoFloat g;
  g.crop_top = g.crop_bottom = g.crop_left = g.crop_right = 1;
  auto wh = g.transform_forward(128, 64);
  if (wh.first != 126 || wh.second != 62) return 3;
  OpTransform t;
  t.rotation = IPB_ROT_90;
  wh = t.transform_forward(128, 64);
  if (wh.first != 64 || wh.second != 128) return 4;
  OpRotateCrop rc;
  wh = rc.transform_forward(100, 50);
  if (wh.first != 100 || wh.second != 50) return 5;
  if (ipb_version() != IPB_VERSION) return 6;
  std::printf("ok\n");
  return 0;
}

static int write_file(const std::string &path, const void *p, size_t n) {
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) return 1;
  const size_t k = std::fwrite(p, 1, n, f);
  std::fclose(f);
  return k == n ? 0 : 1;
}

int main(int argc, char **argv) {
  if (argc >= 2 && !std::strcmp(argv[1], "sizes")) return sizes();
  if (argc < 6 || std::strcmp(argv[1], "run")) {
    std::fprintf(stderr, "usage: host_demo sizes | run W H SEED OUT.bin [MAXWIDTH]\n");
    return 2;
  }
  const size_t W = std::strtoul(argv[2], nullptr, 10), H = std::strtoul(argv[3], nullptr, 10);
  const uint64_t seed = std::strtoull(argv[4], nullptr, 10);
  const std::string out = argv[5];
  const size_t maxwidth = argc > 6 ? std::strtoul(argv[6], nullptr, 10) : 0;
  std::vector<uint16_t> raw(W * H);
  for (size_t i = 0; i < raw.size(); i++) raw[i] = (uint16_t)(splitmix64(seed ^ (uint64_t)i) & 16383u);
  try {
    Context ctx(0);
    auto p = Pipeline::new_from_source(ctx, ImageSource::Raw(raw.data(), W, H));
    fill_ops(p->ops());
    p->settings().maxwidth = maxwidth;
    SRGBImage img = p->output_8bit();
    if (write_file(out, img.data.data(), img.data.size())) return 3;
    std::printf("%zu %zu\n", img.width, img.height);

    // the same frame through the ImageOp objects, one run() per op, in the order of all_ops! (pipeline.rs:211-226)
    PipelineGlobals g{&ctx, ImageSource::Raw(raw.data(), W, H), PipelineSettings()};
    g.settings.maxwidth = maxwidth;
    g.settings.demosaic_width = p->settings().demosaic_width;   // what Pipeline::run negotiated (pipeline.rs:331-338)
    g.settings.demosaic_height = p->settings().demosaic_height;
    OpGoFloat gofloat; OpDemosaic demosaic; OpRotateCrop rotatecrop; OpToLab tolab; OpBaseCurve basecurve; OpFromLab fromlab;
    OpGamma gamma; OpTransform transform;
    PipelineOps mine{gofloat, demosaic, rotatecrop, tolab, basecurve, transform};
    fill_ops(mine);
    std::vector<const ImageOp *> chain = {&gofloat, &demosaic, &rotatecrop, &tolab, &basecurve, &fromlab, &gamma, &transform};
    OpBuffer buf;
    for (const ImageOp *op : chain) buf = op->run(g, buf);
    std::vector<uint8_t> packed(buf.width() * buf.height() * 3);
    ctx.check(ipb_pack_8bit(ctx.handle(), buf.handle(), packed.data(), 0));
    if (write_file(out + ".ops", packed.data(), packed.size())) return 3;
    std::printf("%zu %zu %llu\n", buf.width(), buf.height(), ctx.launch_count());
  } catch (const Error &e) {
    std::fprintf(stderr, "imagepipe error %d: %s\n", e.code, e.what());
    return 4;
  }
  return 0;
}
