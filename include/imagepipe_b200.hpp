// imagepipe_b200.hpp — C++17 host mirror of the reference's Pipeline / ImageOp surface over the C ABI (ipb200.h).
//
// The reference is compiled code (Rust); no Rust toolchain exists in this image, so the compiled host side above the
// C ABI is this header: the same types, names, argument meaning and error behaviour as pedrocr/imagepipe, every
// method a thin call into libipb200.so (no pixel is computed here).  INTEGRATION.md shows the equivalent Rust.
//
//   reference (paths under the reference tree)                       here
//   OpBuffer, Arc<OpBuffer>            src/buffer.rs:4-32             imagepipe::OpBuffer (value type holding one reference)
//   ImageSource::{Raw, Other}          src/pipeline.rs:46-50          imagepipe::ImageSource
//   PipelineSettings, PipelineGlobals  src/pipeline.rs:110-152        same names
//   trait ImageOp                      src/pipeline.rs:82-108         imagepipe::ImageOp (name, run, transform_forward/reverse, reset)
//   OpGoFloat .. OpTransform           src/ops/*.rs                   same names, fields = the C PODs they derive from
//   PipelineOps, Pipeline              src/pipeline.rs:154-469        same names; run / output_8bit / output_16bit
//   PipelineCache                      src/pipeline.rs:43,257-260     imagepipe::PipelineCache
//
// Errors: the reference's ops are infallible by signature and panic on invariant violations; here every failing C call
// throws imagepipe::Error (status code + ipb_last_error text).
#ifndef IMAGEPIPE_B200_HPP
#define IMAGEPIPE_B200_HPP

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ipb200.h"

namespace imagepipe {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

// device + stream + tables; one per thread that runs pipelines (pipelines on different contexts are independent)
class Context {
 public:
  explicit Context(int device = 0, void *stream = nullptr) {
    if (int rc = ipb_ctx_create(device, stream, &h_)) throw Error(rc, ipb_last_error(nullptr));
  }
  ~Context() { ipb_ctx_destroy(h_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  ipb_ctx *handle() const { return h_; }
  void synchronize() const { check(ipb_ctx_synchronize(h_)); }
  unsigned long long launch_count() const { return ipb_ctx_launch_count(h_); }
  void check(int rc) const {
    if (rc != IPB_OK) throw Error(rc, ipb_last_error(h_));
  }

 private:
  ipb_ctx *h_ = nullptr;
};

// OpBuffer (src/buffer.rs:4-11) resident on the device.  Copying the object clones the Arc (same pixels).
class OpBuffer {
 public:
  OpBuffer() = default;
  OpBuffer(const Context &ctx, ipb_buffer *owned) : ctx_(&ctx), h_(owned) {}
  OpBuffer(const OpBuffer &o) : ctx_(o.ctx_), h_(o.h_) { if (h_) ipb_buffer_retain(h_); }
  OpBuffer(OpBuffer &&o) noexcept : ctx_(o.ctx_), h_(o.h_) { o.h_ = nullptr; }
  OpBuffer &operator=(OpBuffer o) { std::swap(ctx_, o.ctx_); std::swap(h_, o.h_); return *this; }
  ~OpBuffer() { if (h_) ipb_buffer_release(h_); }

  // OpBuffer::new (buffer.rs:24-32)
  static OpBuffer create(const Context &ctx, size_t width, size_t height, size_t colors, bool monochrome) {
    ipb_buffer *b = nullptr;
    ctx.check(ipb_buffer_new(ctx.handle(), width, height, colors, monochrome, &b));
    return OpBuffer(ctx, b);
  }
  static OpBuffer from_host(const Context &ctx, size_t width, size_t height, size_t colors, bool monochrome, const float *data) {
    ipb_buffer *b = nullptr;
    ctx.check(ipb_buffer_upload(ctx.handle(), width, height, colors, monochrome, data, &b));
    return OpBuffer(ctx, b);
  }
  size_t width() const { return ipb_buffer_width(h_); }
  size_t height() const { return ipb_buffer_height(h_); }
  size_t colors() const { return ipb_buffer_colors(h_); }
  bool monochrome() const { return ipb_buffer_monochrome(h_) != 0; }
  std::vector<float> data() const {  // the reference's `data: Vec<f32>`, downloaded
    std::vector<float> v(width() * height() * colors());
    ctx_->check(ipb_buffer_download(ctx_->handle(), h_, v.data()));
    return v;
  }
  bool same_arc(const OpBuffer &o) const { return h_ == o.h_; }
  ipb_buffer *handle() const { return h_; }
  const Context &context() const { return *ctx_; }

 private:
  const Context *ctx_ = nullptr;
  ipb_buffer *h_ = nullptr;
};

// ImageSource (pipeline.rs:46-50): the decoded raster stays where the caller keeps it (host or device)
struct ImageSource : ipb_source {
  static ImageSource Raw(const uint16_t *data, size_t width, size_t height, size_t cpp = 1, bool on_device = false) {
    ImageSource s;
    s.kind = IPB_SRC_RAW_U16; s.width = width; s.height = height; s.cpp = cpp; s.data = data; s.on_device = on_device;
    return s;
  }
  static ImageSource RawFloat(const float *data, size_t width, size_t height, size_t cpp = 1, bool on_device = false) {
    ImageSource s;
    s.kind = IPB_SRC_RAW_F32; s.width = width; s.height = height; s.cpp = cpp; s.data = data; s.on_device = on_device;
    return s;
  }
  static ImageSource OtherRgb8(const uint8_t *data, size_t width, size_t height, bool on_device = false) {
    ImageSource s;
    s.kind = IPB_SRC_RGB8; s.width = width; s.height = height; s.cpp = 3; s.data = data; s.on_device = on_device;
    return s;
  }
  static ImageSource OtherRgb16(const uint16_t *data, size_t width, size_t height, bool on_device = false) {
    ImageSource s;
    s.kind = IPB_SRC_RGB16; s.width = width; s.height = height; s.cpp = 3; s.data = data; s.on_device = on_device;
    return s;
  }
};

struct PipelineSettings : ipb_settings {  // pipeline.rs:110-131
  PipelineSettings() : ipb_settings{0, 0, 0, 0, 0, 1} {}
};

struct PipelineGlobals {  // pipeline.rs:139-152
  const Context *ctx;
  ImageSource image;
  PipelineSettings settings;
};

// trait ImageOp (pipeline.rs:82-108)
struct ImageOp {
  virtual ~ImageOp() = default;
  virtual const char *name() const = 0;
  virtual OpBuffer run(const PipelineGlobals &pipeline, const OpBuffer &buf) const = 0;
  virtual std::pair<size_t, size_t> transform_forward(size_t width, size_t height) { return {width, height}; }
  virtual std::pair<size_t, size_t> transform_reverse(size_t width, size_t height) { return {width, height}; }
  virtual void reset() {}
};

namespace detail {
template <class F>
OpBuffer run_op(const PipelineGlobals &g, F &&call) {
  ipb_buffer *out = nullptr;
  g.ctx->check(call(&out));
  return OpBuffer(*g.ctx, out);
}
}  // namespace detail

struct OpGoFloat : ImageOp, ipb_gofloat {  // src/ops/gofloat.rs
  OpGoFloat() : ipb_gofloat{} {}
  const char *name() const override { return "gofloat"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_gofloat_run(g.ctx->handle(), this, &g.image, o); });
  }
  std::pair<size_t, size_t> transform_forward(size_t w, size_t h) override {
    size_t ow, oh;
    ipb_gofloat_transform_forward(this, w, h, &ow, &oh);
    return {ow, oh};
  }
};

struct OpDemosaic : ImageOp, ipb_demosaic {  // src/ops/demosaic.rs
  OpDemosaic() : ipb_demosaic{} {}
  void set_cfa(const std::string &pattern) { std::strncpy(cfa, pattern.c_str(), sizeof(cfa) - 1); }
  const char *name() const override { return "demosaic"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &buf) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_demosaic_run(g.ctx->handle(), this, &g.settings, buf.handle(), o); });
  }
};

struct OpRotateCrop : ImageOp, ipb_rotatecrop {  // src/ops/rotatecrop.rs
  OpRotateCrop() : ipb_rotatecrop{} { input_ratio = 1.0f; }
  const char *name() const override { return "rotatecrop"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &buf) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_rotatecrop_run(g.ctx->handle(), this, buf.handle(), o); });
  }
  std::pair<size_t, size_t> transform_forward(size_t w, size_t h) override {
    size_t ow, oh;
    ipb_rotatecrop_transform_forward(this, w, h, &ow, &oh);
    return {ow, oh};
  }
  std::pair<size_t, size_t> transform_reverse(size_t w, size_t h) override {
    size_t ow, oh;
    ipb_rotatecrop_transform_reverse(this, w, h, &ow, &oh);
    return {ow, oh};
  }
  void reset() override { ipb_rotatecrop_reset(this); }
};

struct OpToLab : ImageOp, ipb_tolab {  // src/ops/colorspaces.rs:5-113
  OpToLab() : ipb_tolab{} {}
  const char *name() const override { return "to_lab"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &buf) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_tolab_run(g.ctx->handle(), this, buf.handle(), o); });
  }
};

struct OpBaseCurve : ImageOp, ipb_basecurve {  // src/ops/curves.rs
  OpBaseCurve() : ipb_basecurve{} {}
  void set_points(const std::vector<std::pair<float, float>> &pts) {
    if (pts.size() > IPB_MAX_CURVE_POINTS) throw Error(IPB_ERR_INVALID, "too many curve points");
    npoints = pts.size();
    for (size_t i = 0; i < pts.size(); i++) { points[i][0] = pts[i].first; points[i][1] = pts[i].second; }
  }
  const char *name() const override { return "basecurve"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &buf) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_basecurve_run(g.ctx->handle(), this, buf.handle(), o); });
  }
};

struct OpFromLab : ImageOp {  // src/ops/colorspaces.rs:115-138
  const char *name() const override { return "from_lab"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &buf) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_fromlab_run(g.ctx->handle(), buf.handle(), o); });
  }
};

struct OpGamma : ImageOp {  // src/ops/gamma.rs
  const char *name() const override { return "gamma"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &buf) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_gamma_run(g.ctx->handle(), &g.settings, buf.handle(), o); });
  }
};

struct OpTransform : ImageOp, ipb_transform {  // src/ops/transform.rs
  OpTransform() : ipb_transform{} {}
  const char *name() const override { return "transform"; }
  OpBuffer run(const PipelineGlobals &g, const OpBuffer &buf) const override {
    return detail::run_op(g, [&](ipb_buffer **o) { return ipb_transform_run(g.ctx->handle(), this, buf.handle(), o); });
  }
  std::pair<size_t, size_t> transform_forward(size_t w, size_t h) override {
    size_t ow, oh;
    ipb_transform_transform_forward(this, w, h, &ow, &oh);
    return {ow, oh};
  }
  std::pair<size_t, size_t> transform_reverse(size_t w, size_t h) override { return transform_forward(w, h); }
};

// PipelineCache = MultiCache<BufHash, OpBuffer> (pipeline.rs:43); Pipeline::new_cache(size) (:257-260)
class PipelineCache {
 public:
  PipelineCache(const Context &ctx, size_t size) { ctx.check(ipb_cache_create(ctx.handle(), size, &h_)); }
  ~PipelineCache() { ipb_cache_destroy(h_); }
  PipelineCache(const PipelineCache &) = delete;
  PipelineCache &operator=(const PipelineCache &) = delete;
  size_t bytes() const { return ipb_cache_bytes(h_); }
  size_t entries() const { return ipb_cache_entries(h_); }
  ipb_cache *handle() const { return h_; }

 private:
  ipb_cache *h_ = nullptr;
};

struct SRGBImage {  // pipeline.rs:26-32
  size_t width, height;
  std::vector<uint8_t> data;
};
struct SRGBImage16 {  // pipeline.rs:34-41
  size_t width, height;
  std::vector<uint16_t> data;
};

// PipelineOps (pipeline.rs:154-164) as live views into the native pipeline's parameter block
struct PipelineOps {
  ipb_gofloat &gofloat;
  ipb_demosaic &demosaic;
  ipb_rotatecrop &rotatecrop;
  ipb_tolab &tolab;
  ipb_basecurve &basecurve;
  ipb_transform &transform;
};

// Pipeline (pipeline.rs:245-470)
class Pipeline {
 public:
  // Pipeline::new_from_source (pipeline.rs:274-284); the raster is not copied and must outlive the pipeline
  static std::unique_ptr<Pipeline> new_from_source(const Context &ctx, const ImageSource &img) {
    return std::unique_ptr<Pipeline>(new Pipeline(ctx, img));
  }
  static std::unique_ptr<PipelineCache> new_cache(const Context &ctx, size_t size) {
    return std::unique_ptr<PipelineCache>(new PipelineCache(ctx, size));
  }
  ~Pipeline() { ipb_pipeline_destroy(h_); }
  Pipeline(const Pipeline &) = delete;
  Pipeline &operator=(const Pipeline &) = delete;

  PipelineOps ops() {
    ipb_ops *o = ipb_pipeline_ops(h_);
    return PipelineOps{o->gofloat, o->demosaic, o->rotatecrop, o->tolab, o->basecurve, o->transform};
  }
  ipb_settings &settings() { return *ipb_pipeline_settings(h_); }  // pipeline.globals.settings

  // Pipeline::run(cache) (pipeline.rs:311-375): 3-channel f32 OpBuffer
  OpBuffer run(const PipelineCache *cache = nullptr) {
    ipb_buffer *out = nullptr;
    ctx_->check(cache ? ipb_pipeline_run_cached(h_, cache->handle(), &out) : ipb_pipeline_run(h_, &out));
    return OpBuffer(*ctx_, out);
  }
  std::pair<size_t, size_t> output_size() {
    size_t w, h;
    ctx_->check(ipb_pipeline_output_size(h_, &w, &h));
    return {w, h};
  }
  // Pipeline::output_8bit / output_16bit (pipeline.rs:377-469); with a cache the library still tries the non-raw fast
  // path first (pipeline.rs:381, :428), then runs from the first changed op and packs
  SRGBImage output_8bit(const PipelineCache *cache = nullptr) {
    SRGBImage img;
    auto wh = output_size();
    img.data.resize(wh.first * wh.second * 3);
    ctx_->check(ipb_pipeline_output_8bit_cached(h_, cache ? cache->handle() : nullptr, img.data.data(), img.data.size(), 0,
                                                &img.width, &img.height));
    img.data.resize(img.width * img.height * 3);
    return img;
  }
  SRGBImage16 output_16bit(const PipelineCache *cache = nullptr) {
    SRGBImage16 img;
    auto wh = output_size();
    img.data.resize(wh.first * wh.second * 3);
    ctx_->check(ipb_pipeline_output_16bit_cached(h_, cache ? cache->handle() : nullptr, img.data.data(), img.data.size(), 0,
                                                 &img.width, &img.height));
    img.data.resize(img.width * img.height * 3);
    return img;
  }
  ipb_pipeline *handle() const { return h_; }

 private:
  Pipeline(const Context &ctx, const ImageSource &img) : ctx_(&ctx) {
    ctx.check(ipb_pipeline_create(ctx.handle(), &img, nullptr, &h_));
  }
  const Context *ctx_;
  ipb_pipeline *h_ = nullptr;
};

}  // namespace imagepipe

#endif  // IMAGEPIPE_B200_HPP
