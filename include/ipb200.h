/*
 * ipb200.h — C ABI of imagepipe-b200 (libipb200.so): the B200-native (sm_100a) replacement for
 * the OpBuffer hot path of pedrocr/imagepipe.
 *
 * This header is the drop-in boundary: it is exactly what a Rust `extern "C"` block in the
 * reference would bind (INTEGRATION.md shows that block and the `impl ImageOp` shims).  Plain
 * pointers and sizes only; no torch / CUDA types in any signature (a CUDA stream is passed as an
 * opaque `void *`).  Every entry point cites the reference interface it replaces; paths are
 * relative to the reference tree (pedrocr/imagepipe @ 65ca96ce).
 *
 * Conventions
 *  - every function that can fail returns an ipb_status (0 == IPB_OK); the message for the last
 *    failure on a context is ipb_last_error(ctx).  Nothing aborts the host process (the reference
 *    panics on invariant violations: scaling.rs:133,148, transform.rs:88,98).
 *  - calls on one ipb_ctx are ordered on its CUDA stream; entry points that return host data
 *    synchronise before returning, everything else is asynchronous — call ipb_ctx_synchronize().
 *    Distinct contexts are independent and may be used from different threads.
 *  - there is NO CPU fallback: every op runs a CUDA kernel or returns IPB_ERR_CUDA.
 */
#ifndef IPB200_H
#define IPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define IPB_VERSION 100 /* 0.1.0 */

typedef enum ipb_status {
  IPB_OK = 0,
  IPB_ERR_INVALID = 1,     /* bad argument / null pointer / size mismatch */
  IPB_ERR_BAD_COLORS = 2,  /* op given a buffer with the wrong channel count (reference: assert_eq! panic) */
  IPB_ERR_BAD_CFA = 3,     /* CFA pattern string rawloader::CFA::new would panic on */
  IPB_ERR_CUDA = 4,        /* CUDA runtime / driver error, or no usable device */
  IPB_ERR_UNSUPPORTED = 5, /* combination outside the hot path */
  IPB_ERR_NOMEM = 6
} ipb_status;

typedef struct ipb_ctx ipb_ctx;           /* device + stream + uploaded LUTs/constants */
typedef struct ipb_buffer ipb_buffer;     /* device-resident OpBuffer (src/buffer.rs:4-11), ref-counted like Arc<OpBuffer> */
typedef struct ipb_pipeline ipb_pipeline; /* Pipeline (src/pipeline.rs:245-249) */
typedef struct ipb_cache ipb_cache;       /* PipelineCache = MultiCache<BufHash, OpBuffer> (src/pipeline.rs:43), device-resident */

/* ------------------------------------------------------------------ parameter PODs
 * 1:1 with the reference's serde structs so the Rust side can fill them field by field. */

/* OpGoFloat — src/ops/gofloat.rs:4-12 */
typedef struct ipb_gofloat {
  size_t crop_top, crop_right, crop_bottom, crop_left;
  int is_cfa;
  float blacklevels[4];
  float whitelevels[4];
} ipb_gofloat;

/* OpDemosaic — src/ops/demosaic.rs:4-6; `cfa` is the rawloader CFA pattern string (≤144 chars, NUL terminated) */
typedef struct ipb_demosaic {
  char cfa[148];
} ipb_demosaic;

/* OpRotateCrop — src/ops/rotatecrop.rs:10-18 (input_ratio/output_size are the op's private state) */
typedef struct ipb_rotatecrop {
  float crop_top, crop_right, crop_bottom, crop_left, rotation;
  float input_ratio;
  int has_output_size;
  size_t output_width, output_height;
} ipb_rotatecrop;

/* OpToLab — src/ops/colorspaces.rs:5-10 */
typedef struct ipb_tolab {
  float cam_to_xyz[3][4];
  float cam_to_xyz_normalized[3][4];
  float xyz_to_cam[4][3];
  float wb_coeffs[4];
} ipb_tolab;

/* OpBaseCurve — src/ops/curves.rs:6-9 (points: Vec<(f32,f32)>, capped at 32) */
#define IPB_MAX_CURVE_POINTS 32
typedef struct ipb_basecurve {
  float exposure;
  size_t npoints;
  float points[IPB_MAX_CURVE_POINTS][2];
} ipb_basecurve;

/* OpTransform / Rotation — src/ops/transform.rs:6-19 */
enum { IPB_ROT_NORMAL = 0, IPB_ROT_90 = 1, IPB_ROT_180 = 2, IPB_ROT_270 = 3 };
typedef struct ipb_transform {
  int rotation;
  int fliph, flipv;
} ipb_transform;

/* PipelineSettings — src/pipeline.rs:110-118 */
typedef struct ipb_settings {
  size_t maxwidth, maxheight, demosaic_width, demosaic_height;
  int linear;
  int use_fastpath;
} ipb_settings;

/* ImageSource — src/pipeline.rs:46-50.  RAW_* is rawloader::RawImage{width,height,cpp,data}
 * (Integer = u16 / Float = f32), RGB8/RGB16 is the `image` crate's to_rgb8()/to_rgb16() raster.
 * `data` may live on the host (pageable or pinned) or, with on_device != 0, on the context's GPU. */
enum { IPB_SRC_RAW_U16 = 0, IPB_SRC_RAW_F32 = 1, IPB_SRC_RGB8 = 2, IPB_SRC_RGB16 = 3 };
typedef struct ipb_source {
  int kind;
  size_t width, height, cpp;
  const void *data;
  int on_device;
} ipb_source;

/* PipelineOps — src/pipeline.rs:154-164 (OpFromLab and OpGamma have no fields) */
typedef struct ipb_ops {
  ipb_gofloat gofloat;
  ipb_demosaic demosaic;
  ipb_rotatecrop rotatecrop;
  ipb_tolab tolab;
  ipb_basecurve basecurve;
  ipb_transform transform;
} ipb_ops;

/* Row stripe of a frame for multi-GPU sharding (no reference equivalent).
 * The source passed with a stripe holds only rows [src_row0, src_row0 + source.height) of a
 * full frame `full_height` rows tall; the call produces output rows [out_row0, out_row1) of the
 * full-frame result.  Use ipb_pipeline_stripe_rows() to learn which source rows a stripe needs. */
typedef struct ipb_stripe {
  size_t full_height;
  size_t src_row0;
  size_t out_row0, out_row1;
} ipb_stripe;

/* ------------------------------------------------------------------ context */

int ipb_version(void);
/* stream: a cudaStream_t to run on (e.g. torch's current stream) or NULL for a private stream. */
int ipb_ctx_create(int device, void *stream, ipb_ctx **out);
void ipb_ctx_destroy(ipb_ctx *ctx);
int ipb_ctx_set_stream(ipb_ctx *ctx, void *stream);
int ipb_ctx_synchronize(ipb_ctx *ctx);
const char *ipb_last_error(const ipb_ctx *ctx); /* ctx == NULL: last ipb_ctx_create failure of this thread */
/* number of CUDA kernels this context has launched so far (bench.py's gpu_launches) */
unsigned long long ipb_ctx_launch_count(const ipb_ctx *ctx);
/* pinned host memory for the host<->device paths (plain cudaHostAlloc/cudaFreeHost) */
int ipb_host_alloc(size_t bytes, void **out);
void ipb_host_free(void *p);

/* plain device memory for sources and destinations that live on the GPU (stream-ordered on the context's
 * stream; upload/download synchronise before returning) */
int ipb_device_alloc(ipb_ctx *ctx, size_t bytes, void **out);
int ipb_device_free(ipb_ctx *ctx, void *dptr);
int ipb_device_upload(ipb_ctx *ctx, void *dptr, const void *host, size_t bytes);
int ipb_device_download(ipb_ctx *ctx, void *host, const void *dptr, size_t bytes);

/* ------------------------------------------------------------------ OpBuffer — src/buffer.rs */

/* OpBuffer::new (buffer.rs:24-32): zero-filled device buffer */
int ipb_buffer_new(ipb_ctx *ctx, size_t width, size_t height, size_t colors, int monochrome, ipb_buffer **out);
/* host Vec<f32> -> device OpBuffer (same interleaved row-major layout, buffer.rs:4-11) */
int ipb_buffer_upload(ipb_ctx *ctx, size_t width, size_t height, size_t colors, int monochrome,
                      const float *host, ipb_buffer **out);
/* wrap caller-owned device memory (not freed on release) */
int ipb_buffer_wrap(ipb_ctx *ctx, size_t width, size_t height, size_t colors, int monochrome, void *dptr,
                    ipb_buffer **out);
/* device OpBuffer -> host Vec<f32>; synchronises */
int ipb_buffer_download(ipb_ctx *ctx, const ipb_buffer *buf, float *host);
void ipb_buffer_retain(ipb_buffer *buf);  /* Arc::clone */
void ipb_buffer_release(ipb_buffer *buf); /* drop(Arc) */
size_t ipb_buffer_width(const ipb_buffer *buf);
size_t ipb_buffer_height(const ipb_buffer *buf);
size_t ipb_buffer_colors(const ipb_buffer *buf);
int ipb_buffer_monochrome(const ipb_buffer *buf);
void *ipb_buffer_device_ptr(const ipb_buffer *buf);

/* ------------------------------------------------------------------ ImageOp::run — src/pipeline.rs:82-108
 * One entry per op of PipelineOps, same order as all_ops! (pipeline.rs:211-226).  `*out` receives a
 * new reference; ops that pass their input through (the reference returns the same Arc) return `in`
 * retained, so callers always release `*out` exactly once. */

/* OpGoFloat::run — gofloat.rs:50-62 (run_raw :84-169, run_other :171-201) */
int ipb_gofloat_run(ipb_ctx *ctx, const ipb_gofloat *op, const ipb_source *image, ipb_buffer **out);
/* OpDemosaic::run — demosaic.rs:27-61 (full() :67-119; scaled_demosaic / scale_down_opbuf scaling.rs:132-160) */
int ipb_demosaic_run(ipb_ctx *ctx, const ipb_demosaic *op, const ipb_settings *settings, ipb_buffer *in,
                     ipb_buffer **out);
/* OpRotateCrop::run — rotatecrop.rs:39-64 (OpBuffer::transform buffer.rs:62-79) */
int ipb_rotatecrop_run(ipb_ctx *ctx, const ipb_rotatecrop *op, ipb_buffer *in, ipb_buffer **out);
/* OpToLab::run — colorspaces.rs:89-112 */
int ipb_tolab_run(ipb_ctx *ctx, const ipb_tolab *op, ipb_buffer *in, ipb_buffer **out);
/* OpBaseCurve::run — curves.rs:33-49 */
int ipb_basecurve_run(ipb_ctx *ctx, const ipb_basecurve *op, ipb_buffer *in, ipb_buffer **out);
/* OpFromLab::run — colorspaces.rs:127-137 */
int ipb_fromlab_run(ipb_ctx *ctx, ipb_buffer *in, ipb_buffer **out);
/* OpGamma::run — gamma.rs:16-26 */
int ipb_gamma_run(ipb_ctx *ctx, const ipb_settings *settings, ipb_buffer *in, ipb_buffer **out);
/* OpTransform::run — transform.rs:56-73 (rotate_buffer :87-144); bit-exact data movement */
int ipb_transform_run(ipb_ctx *ctx, const ipb_transform *op, ipb_buffer *in, ipb_buffer **out);

/* ImageOp::transform_forward / transform_reverse / reset — host-only size negotiation
 * (gofloat.rs:64-82, rotatecrop.rs:66-86,111-163, transform.rs:75-84) and scaling.rs:8-32 */
void ipb_gofloat_transform_forward(const ipb_gofloat *op, size_t w, size_t h, size_t *ow, size_t *oh);
void ipb_rotatecrop_transform_forward(ipb_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh);
void ipb_rotatecrop_transform_reverse(ipb_rotatecrop *op, size_t w, size_t h, size_t *ow, size_t *oh);
void ipb_rotatecrop_reset(ipb_rotatecrop *op);
void ipb_transform_transform_forward(const ipb_transform *op, size_t w, size_t h, size_t *ow, size_t *oh);
void ipb_scaling_size(size_t w, size_t h, size_t maxw, size_t maxh, size_t *ow, size_t *oh);
float ipb_calculate_scale(size_t w, size_t h, size_t maxw, size_t maxh);

/* SplineFunc::new(&op->points).interpolate(v) for n host values — curves.rs:68-157 (OpBaseCurve::get_spline,
 * :53-55: the exposure field is not applied).  Coefficients are built on the host, evaluated by a kernel. */
int ipb_spline_eval(ipb_ctx *ctx, const ipb_basecurve *op, const float *in, float *out, size_t n);

/* output8bit / output16bit pack loops — pipeline.rs:408-414,455-461, color_conversions.rs:323-330.
 * dst holds width*height*3 elements, on the host (dst_on_device == 0; synchronises) or the device. */
int ipb_pack_8bit(ipb_ctx *ctx, const ipb_buffer *in, uint8_t *dst, int dst_on_device);
int ipb_pack_16bit(ipb_ctx *ctx, const ipb_buffer *in, uint16_t *dst, int dst_on_device);

/* scale_down_srgb / scale_down_srgb16 — scaling.rs:162-182 (the non-raw fast path's resampler) */
int ipb_scale_down_srgb(ipb_ctx *ctx, const uint8_t *src, size_t w, size_t h, size_t nw, size_t nh, uint8_t *dst,
                        int on_device);
int ipb_scale_down_srgb16(ipb_ctx *ctx, const uint16_t *src, size_t w, size_t h, size_t nw, size_t nh,
                          uint16_t *dst, int on_device);

/* EXTENSION — Lanczos-a separable resampler of an f32 OpBuffer (any channel count), 1 <= a <= 8 (3 is the usual
 * choice).  The reference has NO Lanczos resampler: src/scaling.rs:101-103 is a FIXME that names it as a possible
 * replacement of the paraboloid window.  Never used by the pipeline entry points; parity is against
 * oracle/lanczos.c, not the reference. */
int ipb_lanczos_resize(ipb_ctx *ctx, ipb_buffer *in, size_t nwidth, size_t nheight, int a, ipb_buffer **out);

/* ------------------------------------------------------------------ Pipeline — src/pipeline.rs:257-470 */

/* PipelineOps::new for everything that does not need rawloader metadata (pipeline.rs:166-179):
 * raw sources get basecurve [(0.5,0.6)], is_cfa=1, wb 1.0; other sources get the sRGB matrices. */
void ipb_ops_default(ipb_ops *ops, const ipb_source *image);
/* Pipeline::new_from_source (pipeline.rs:274-284).  ops == NULL: ipb_ops_default(). The source's pixel
 * data is NOT copied: it must stay valid until the pipeline is destroyed or the source replaced. */
int ipb_pipeline_create(ipb_ctx *ctx, const ipb_source *image, const ipb_ops *ops, ipb_pipeline **out);
void ipb_pipeline_destroy(ipb_pipeline *p);
/* pipeline.ops / pipeline.globals.settings are public fields in the reference; these return mutable views */
ipb_ops *ipb_pipeline_ops(ipb_pipeline *p);
ipb_settings *ipb_pipeline_settings(ipb_pipeline *p);
/* Swap the source (the next frame of a batch).  A host-resident source is copied to the device asynchronously on the
 * context's stream by the calls that read it: when their result stays on the device (ipb_pipeline_run, *_run entry
 * points, dst_on_device != 0) the host pixels must stay untouched until ipb_ctx_synchronize (or any call that
 * returns host data) — the lifetime rule of every asynchronous copy from pinned memory. */
int ipb_pipeline_set_source(ipb_pipeline *p, const ipb_source *image);
/* 1 (default): Pipeline::run may use the fused raw->sRGB kernel when the op chain allows it;
 * 0: always run op by op (one kernel + one OpBuffer per op, like the reference). */
int ipb_pipeline_set_fused(ipb_pipeline *p, int fused);
/* 1 (default): the full-resolution fused kernel stages raw tiles with TMA when the source allows it (16-byte
 * aligned base and row pitch); 0: always use the plain-load staging path.  Results are identical. */
int ipb_pipeline_set_tma(ipb_pipeline *p, int use_tma);
/* 8-bit output of a full-resolution RGB Bayer frame (the path of Pipeline::output_8bit, pipeline.rs:377-422, for raw
 * sources) runs by default through the speculative kernel: every pixel in cheap contracted arithmetic whose distance
 * from the reference's f32 result is bounded by a certified delta, and the pixels that come within delta of one of
 * the 255 thresholds of output8bit(apply_srgb_gamma(v)) recomputed in the reference's exact arithmetic.  The bytes are
 * identical to the exact kernel's by construction; 0 selects the exact kernel for every pixel. */
int ipb_pipeline_set_speculative(ipb_pipeline *p, int on);
/* Tuning / test hooks of the speculative kernel.  delta > 0 forces the bound used for the threshold test (a huge
 * delta recomputes almost every pixel: results must not change; a delta below the certified one voids the guarantee),
 * 0 restores the certified bound; threads = 512 or 1024 selects the CTA size. */
int ipb_ctx_set_spec(ipb_ctx *ctx, float delta, int threads);
/* counters since the last reset: out[0] = pixels recomputed exactly, out[1] = 1 if a launch found the shared window
 * where the tables' addressing does not expect it (never), out[2] = the certified delta of the current tables as
 * float bits, out[3] = measured max relative error of the XU-pipe cube root as float bits */
int ipb_ctx_spec_stats(ipb_ctx *ctx, unsigned long long out[4], int reset);
/* Measures, on the pipeline's own source frame and parameters, the largest and the mean |cheap - exact| over the
 * clamped linear channel values of every interior pixel, next to the certified bound the kernel would use.  The
 * parity suite asserts max_dev <= delta on every frame it runs.  IPB_ERR_UNSUPPORTED when the pipeline would not take
 * the speculative path. */
int ipb_pipeline_spec_probe(ipb_pipeline *p, float *max_dev, double *mean_dev, float *delta);
/* The certified bound itself, without a context or a GPU (pure host arithmetic, ipb_spec_host.cu): delta[0] = the largest,
 * delta[1..3] = per output channel (R, G, B), for the colour
 * parameters of `ops`, given the relative error of the XU-pipe cube root (ipb_ctx_spec_stats out[3] on a GPU box; the
 * hardware documentation's 2^-22 otherwise).  IPB_ERR_UNSUPPORTED when these parameters would run on the exact kernel. */
int ipb_spec_bound(const ipb_ops *ops, float mufu_rel_err, float delta[4]);
/* Diagnostic, host only: what the speculative kernels' 8-bit stage is built from for a parameter set — the fixed-point
 * gamma table (8192 words), the 255 thresholds of output8bit(apply_srgb_gamma(v)) it encodes, and the constants of the
 * certificate (ipb_spec.h SpecParams: one[], wmul[], amb_t) — so that table and certificate can be checked against the
 * thresholds without a GPU (tests/test_spec_bound_cpu.py).  delta_override > 0 forces the bound. */
int ipb_spec_tables(const ipb_ops *ops, float mufu_rel_err, float delta_override, uint32_t *g8a, float *thresholds,
                    float one[3], uint32_t wmul[3], uint32_t *amb_t, float delta[4]);
/* Diagnostic, host only: 1 when the scaled kernels may divide window coordinates by the skip (scaling.rs:94-98) in the
 * three-instruction reciprocal form for this geometry (cropped frame width x height -> nwidth x nheight) — decided by
 * walking every tap of every window and comparing with IEEE division — else 0 (they then divide the IEEE way). */
int ipb_scaled_division_check(size_t width, size_t height, size_t nwidth, size_t nheight);
/* Host-resident source and/or destination: output_8bit / output_16bit cut the frame into bands of about `megabytes`
 * of PCIe traffic and overlap the H2D copy, the kernel and the D2H copy of neighbouring bands on three streams
 * (default 16: lowest latency of a single call).  0 = one band: whole-frame copies, which use the PCIe link better
 * when several host threads (each with its own context) keep both directions busy across frames.  Same results. */
int ipb_pipeline_set_band_mb(ipb_pipeline *p, int megabytes);
/* size walk of Pipeline::run (pipeline.rs:313-338): final output size; also sets settings.demosaic_* */
int ipb_pipeline_output_size(ipb_pipeline *p, size_t *width, size_t *height);
/* Pipeline::run(None) — pipeline.rs:311-375; result is a 3-channel f32 OpBuffer */
int ipb_pipeline_run(ipb_pipeline *p, ipb_buffer **out);
/* Pipeline::new_cache(size) — pipeline.rs:257-260: a size-bounded (bytes of f32 data, as the reference counts them,
 * :369) least-recently-used cache of device OpBuffers.  One cache may serve several pipelines and threads of the same
 * context. */
int ipb_cache_create(ipb_ctx *ctx, size_t max_bytes, ipb_cache **out);
void ipb_cache_destroy(ipb_cache *cache);
void ipb_cache_clear(ipb_cache *cache);
size_t ipb_cache_bytes(const ipb_cache *cache);
size_t ipb_cache_entries(const ipb_cache *cache);
/* Pipeline::run(Some(&cache)) — pipeline.rs:340-372: the cumulative hash chain over settings and op parameters finds
 * the latest op whose output is cached; the ops after it run one kernel each and every result goes into the cache, so
 * that a parameter change re-runs the pipeline from the first op it affects.  cache == NULL: ipb_pipeline_run. */
int ipb_pipeline_run_cached(ipb_pipeline *p, ipb_cache *cache, ipb_buffer **out);
/* what the last ipb_pipeline_run_cached did: index of the first op executed (0 = gofloat ... 7 = transform, 8 = the
 * final buffer came from the cache) and how many ops ran */
void ipb_pipeline_last_run_info(const ipb_pipeline *p, int *startpos, int *ops_run);
/* Pipeline::output_8bit / output_16bit — pipeline.rs:377-469 (incl. the non-raw fast path).
 * dst capacity is in elements; *width and *height receive the image size. dst_on_device == 0 synchronises. */
int ipb_pipeline_output_8bit(ipb_pipeline *p, uint8_t *dst, size_t dst_capacity, int dst_on_device, size_t *width,
                             size_t *height);
int ipb_pipeline_output_16bit(ipb_pipeline *p, uint16_t *dst, size_t dst_capacity, int dst_on_device,
                              size_t *width, size_t *height);
/* output_8bit(Some(&cache)) / output_16bit(Some(&cache)): like the reference, the non-raw fast path is tried first with
 * or without a cache (pipeline.rs:381, :428); otherwise settings.linear is set, the cached run restarts at the first op
 * whose parameters changed, and the result is packed.  cache == NULL: the calls above. */
int ipb_pipeline_output_8bit_cached(ipb_pipeline *p, ipb_cache *cache, uint8_t *dst, size_t dst_capacity, int dst_on_device,
                                    size_t *width, size_t *height);
int ipb_pipeline_output_16bit_cached(ipb_pipeline *p, ipb_cache *cache, uint16_t *dst, size_t dst_capacity, int dst_on_device,
                                     size_t *width, size_t *height);

/* Row-stripe sharding (multi-GPU).  Only for the fused raw CFA path with a
 * Normal orientation and no rotatecrop. */
/* which source rows [*src_row0, *src_row1) are needed to produce output rows [out_row0, out_row1) */
int ipb_pipeline_stripe_rows(ipb_pipeline *p, size_t out_row0, size_t out_row1, size_t *src_row0, size_t *src_row1);
/* The same question without a context or a GPU (pure host arithmetic: the size walk of pipeline.rs:313-338, the
 * demosaic branch of demosaic.rs:41-60 and the window rows of scaling.rs:69-87): for a `width` x `height` u16 CFA
 * source, *out_width / *out_height receive the size of the result and, when out_row0 < out_row1, [*src_row0, *src_row1)
 * the source rows that output rows [out_row0, out_row1) need.  IPB_ERR_UNSUPPORTED when the chain cannot be sharded by
 * rows (not the fused CFA path, or an orientation other than Normal).  Used by the sharding driver to size halos. */
int ipb_stripe_plan(const ipb_ops *ops, const ipb_settings *settings, size_t width, size_t height, size_t out_row0,
                    size_t out_row1, size_t *src_row0, size_t *src_row1, size_t *out_width, size_t *out_height);
/* like output_8bit but the pipeline's source holds only the rows named by `stripe`; the pipeline must have
 * been created with image.height == stripe->full_height semantics via ipb_pipeline_set_stripe_source(). */
int ipb_pipeline_set_stripe_source(ipb_pipeline *p, const ipb_source *rows, const ipb_stripe *stripe);
int ipb_pipeline_output_8bit_stripe(ipb_pipeline *p, uint8_t *dst, size_t dst_capacity, int dst_on_device,
                                    size_t *width, size_t *rows);
/* A batch of frames in one call (extension: the reference's Pipeline::output_8bit, pipeline.rs:377-422, takes one image;
 * a caller that converts many frames of one camera — BASELINE config 4, or the stripes of several frames in flight —
 * pays kernel start-up and tail once per batch instead of once per frame).  nframes frames of identical geometry and
 * parameters, all on the device: frame k's source rows begin src_stride_rows * k rows after the pipeline's source
 * (the whole image, or the stripe rows given to ipb_pipeline_set_stripe_source), its result dst_stride_bytes * k bytes
 * after dst.  Same bytes as nframes calls of ipb_pipeline_output_8bit / _stripe on shifted pointers.  Only for the
 * fused raw CFA path with Normal orientation (IPB_ERR_UNSUPPORTED otherwise). */
int ipb_pipeline_output_8bit_batch(ipb_pipeline *p, size_t nframes, size_t src_stride_rows, uint8_t *dst,
                                   size_t dst_stride_bytes, size_t dst_capacity, size_t *width, size_t *rows);

/* Halo exchange between stripe neighbours: the only collective of the path.  One process per
 * GPU; rank r holds stripe r.  NCCL is resolved at run time (dlopen of libnccl.so.2): the library links neither NCCL
 * nor any framework.  Rank 0 draws a unique id, the host program hands its bytes to every rank through whatever channel
 * it has (torch.distributed in bench.py, MPI, a file), every rank creates its communicator on its device and stream.
 * ipb_halo_exchange enqueues, on that stream, one grouped ncclSend / ncclRecv per neighbour and buffer: the rows this
 * rank owns and a neighbour's stencil reads (1 raw row for demosaic::full, demosaic.rs:70-74; the window rows of
 * scaled_demosaic, scaling.rs:77-87) go out, the rows this rank's stencil reads come in.  `bufs`: nbufs device
 * buffers of identical layout (several frames may share one group, which amortises NCCL's launch latency); offsets
 * are bytes from the start of each buffer.  Stream-ordered: no host synchronisation; capturable into a CUDA graph. */
#define IPB_COMM_ID_BYTES 128
typedef struct ipb_comm ipb_comm;
typedef struct ipb_halo {
  size_t send_up_off, send_up_bytes;     /* rows sent to rank - 1 (0 bytes: nothing; must be 0 on rank 0) */
  size_t recv_up_off, recv_up_bytes;     /* rows received from rank - 1 */
  size_t send_down_off, send_down_bytes; /* rows sent to rank + 1 (must be 0 on the last rank) */
  size_t recv_down_off, recv_down_bytes; /* rows received from rank + 1 */
} ipb_halo;
int ipb_comm_unique_id(unsigned char id[IPB_COMM_ID_BYTES]);
int ipb_comm_create(int device, void *stream, const unsigned char id[IPB_COMM_ID_BYTES], int rank, int nranks,
                    ipb_comm **out);
void ipb_comm_destroy(ipb_comm *comm);
int ipb_comm_rank(const ipb_comm *comm);
int ipb_comm_size(const ipb_comm *comm);
int ipb_comm_nccl_version(int *version);
const char *ipb_comm_last_error(const ipb_comm *comm);
int ipb_halo_exchange(ipb_comm *comm, void *const *bufs, size_t nbufs, const ipb_halo *halo);

/* ------------------------------------------------------------------ self-checks of derived tables
 * The fused 8-bit path evaluates output8bit(apply_srgb_gamma(v)) (gamma.rs:21, color_conversions.rs:323-325)
 * through a per-segment threshold table derived from SRGB_GAMMA_TRANSFORM.  ipb_selftest_gamma8 compares it on
 * the device with the plain lerp + quantise code for every f32 in [0,1] and a sample of all other bit patterns;
 * ipb_gamma_pack_8bit runs it on n host floats (tests compare with the oracle). */
int ipb_selftest_gamma8(ipb_ctx *ctx, unsigned long long *mismatches);
int ipb_gamma_pack_8bit(ipb_ctx *ctx, const float *in, size_t n, uint8_t *out);

/* ------------------------------------------------------------------ synthetic input (bench/tests)
 * v(i) = splitmix64(seed ^ i) mod 16384 for the pixel with linear index i = row*width + col of the
 * full frame; writes rows [row0, row0+rows) to device memory dptr. */
int ipb_synth_cfa_u16(ipb_ctx *ctx, uint64_t seed, size_t width, size_t row0, size_t rows, uint16_t *dptr);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* IPB200_H */
