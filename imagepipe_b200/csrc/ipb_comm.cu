// ipb_comm.cu — halo exchange of the row-stripe path (BASELINE config 5) inside the C ABI, on NCCL.
//
// One large frame is cut into row stripes, one per GPU (one process per GPU).  demosaic::full needs one raw row above
// and below a stripe (demosaic.rs:70-74), scaled_demosaic the rows of its windows (scaling.rs:77-87): those rows are
// the only data that travels — grouped ncclSend / ncclRecv between stripe neighbours, enqueued on the context's
// stream (stream-ordered with the kernels; capturable into a CUDA graph).  libnccl is resolved at run time
// (dlopen("libnccl.so.2")): a process that already carries an NCCL — e.g. the one PyTorch ships — shares that copy,
// any other host gets the system's; libipb200.so itself links neither NCCL nor torch.
//
// Several buffers in one call (the frames in flight of a step) travel packed: one kernel gathers their outgoing rows into
// a staging area of the communicator, one ncclSend / ncclRecv per neighbour moves all of them, one kernel scatters the
// incoming rows — a message per neighbour instead of one per neighbour and frame (NCCL's cost per small message is what
// the exchange of a step consists of).
#include "../../include/ipb200.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string err;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.err = std::string("libnccl not found: ") + dlerror();
      return;
    }
#define IPB_SYM(field, sym)                                                   \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
  if (!api.field) api.err = std::string("libnccl lacks ") + sym;
    IPB_SYM(GetUniqueId, "ncclGetUniqueId")
    IPB_SYM(CommInitRank, "ncclCommInitRank")
    IPB_SYM(CommDestroy, "ncclCommDestroy")
    IPB_SYM(GroupStart, "ncclGroupStart")
    IPB_SYM(GroupEnd, "ncclGroupEnd")
    IPB_SYM(Send, "ncclSend")
    IPB_SYM(Recv, "ncclRecv")
    IPB_SYM(GetErrorString, "ncclGetErrorString")
    IPB_SYM(GetVersion, "ncclGetVersion")
#undef IPB_SYM
  });
  return api;
}

thread_local std::string g_comm_err;

}  // namespace

struct ipb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
  void *stream = nullptr;
  std::string err;
  unsigned char *stage = nullptr;   // packed exchange: [send up | send down | recv up | recv down], each nbufs rows
  size_t stage_bytes = 0;
};

namespace {

constexpr int kPackMax = 24;   // buffer pointers per pack / unpack launch (passed by value)
struct PackArgs {
  unsigned char *buf[kPackMax];
  int n;
  size_t off_a, bytes_a, off_b, bytes_b;   // region a / b inside every buffer
  unsigned char *stage_a, *stage_b;        // packed rows of region a / b (buffer k at k * bytes)
};
// blockIdx.y = buffer, blockIdx.z = region; 16-byte words when everything is aligned, bytes otherwise
template <bool PACK>
__global__ void k_halo_pack(const PackArgs a) {
  unsigned char *b = a.buf[blockIdx.y];
  const bool second = blockIdx.z != 0;
  const size_t bytes = second ? a.bytes_b : a.bytes_a;
  unsigned char *in_buf = b + (second ? a.off_b : a.off_a);
  unsigned char *in_stage = (second ? a.stage_b : a.stage_a) + (size_t)blockIdx.y * bytes;
  unsigned char *dst = PACK ? in_stage : in_buf;
  const unsigned char *src = PACK ? in_buf : in_stage;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15) == 0) {
    for (size_t i = tid; i < bytes / 16; i += nth) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
  } else {
    for (size_t i = tid; i < bytes; i += nth) dst[i] = src[i];
  }
}

}  // namespace

static int comm_fail(ipb_comm *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  g_comm_err = msg;
  return code;
}

static_assert(sizeof(ncclUniqueId) == IPB_COMM_ID_BYTES, "ipb200.h: IPB_COMM_ID_BYTES is NCCL's unique id size");

extern "C" {

const char *ipb_comm_last_error(const ipb_comm *c) { return c ? c->err.c_str() : g_comm_err.c_str(); }

int ipb_comm_nccl_version(int *version) {
  NcclApi &n = nccl();
  if (!n.err.empty()) return comm_fail(nullptr, IPB_ERR_UNSUPPORTED, n.err);
  return n.GetVersion(version) == ncclSuccess ? IPB_OK : IPB_ERR_CUDA;
}

int ipb_comm_unique_id(unsigned char id[IPB_COMM_ID_BYTES]) {
  if (!id) return IPB_ERR_INVALID;
  NcclApi &n = nccl();
  if (!n.err.empty()) return comm_fail(nullptr, IPB_ERR_UNSUPPORTED, n.err);
  ncclUniqueId u;
  ncclResult_t r = n.GetUniqueId(&u);
  if (r != ncclSuccess) return comm_fail(nullptr, IPB_ERR_CUDA, std::string("ncclGetUniqueId: ") + n.GetErrorString(r));
  memcpy(id, &u, sizeof(u));
  return IPB_OK;
}

int ipb_comm_create(int device, void *stream, const unsigned char id[IPB_COMM_ID_BYTES], int rank, int nranks,
                    ipb_comm **out) {
  if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return IPB_ERR_INVALID;
  *out = nullptr;
  NcclApi &n = nccl();
  if (!n.err.empty()) return comm_fail(nullptr, IPB_ERR_UNSUPPORTED, n.err);
  ipb_comm *c = new (std::nothrow) ipb_comm();
  if (!c) return IPB_ERR_NOMEM;
  c->rank = rank; c->nranks = nranks; c->device = device; c->stream = stream;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete c; return comm_fail(nullptr, IPB_ERR_CUDA, cudaGetErrorString(e)); }
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclResult_t r = n.CommInitRank(&c->comm, nranks, u, rank);
  if (r != ncclSuccess) {
    delete c;
    return comm_fail(nullptr, IPB_ERR_CUDA, std::string("ncclCommInitRank: ") + n.GetErrorString(r));
  }
  *out = c;
  return IPB_OK;
}

void ipb_comm_destroy(ipb_comm *c) {
  if (!c) return;
  if (c->stage) cudaFree(c->stage);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

int ipb_comm_rank(const ipb_comm *c) { return c ? c->rank : -1; }
int ipb_comm_size(const ipb_comm *c) { return c ? c->nranks : 0; }

int ipb_halo_exchange(ipb_comm *c, void *const *bufs, size_t nbufs, const ipb_halo *h) {
  if (!c || !h || (nbufs && !bufs)) return IPB_ERR_INVALID;
  NcclApi &n = nccl();
  const bool up = c->rank > 0, down = c->rank + 1 < c->nranks;
  if ((!up && (h->send_up_bytes || h->recv_up_bytes)) || (!down && (h->send_down_bytes || h->recv_down_bytes)))
    return comm_fail(c, IPB_ERR_INVALID, "halo plan names a neighbour this rank does not have");
  cudaStream_t s = (cudaStream_t)c->stream;

  // ---- packed: all buffers' rows in one message per neighbour and direction
  const size_t per_buf = h->send_up_bytes + h->send_down_bytes + h->recv_up_bytes + h->recv_down_bytes;
  bool packed = nbufs > 1 && per_buf > 0;
  if (packed && c->stage_bytes < nbufs * per_buf) {
    // the staging area grows outside stream capture only (an allocation cannot be captured): a first call inside a
    // capture takes the unpacked path
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      packed = false;
    } else {
      if (c->stage) { cudaStreamSynchronize(s); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
      if (cudaMalloc((void **)&c->stage, nbufs * per_buf + 64) != cudaSuccess) { cudaGetLastError(); packed = false; }
      else c->stage_bytes = nbufs * per_buf;
    }
  }
  if (packed) {
    // offsets rounded so that every region starts on a 16-byte boundary when the row sizes are multiples of 16
    unsigned char *su = c->stage, *sd = su + nbufs * h->send_up_bytes, *ru = sd + nbufs * h->send_down_bytes,
                  *rd = ru + nbufs * h->recv_up_bytes;
    auto run = [&](bool pack) -> cudaError_t {
      for (size_t i0 = 0; i0 < nbufs; i0 += kPackMax) {
        PackArgs a;
        a.n = (int)(nbufs - i0 < (size_t)kPackMax ? nbufs - i0 : (size_t)kPackMax);
        for (int k = 0; k < a.n; k++) a.buf[k] = static_cast<unsigned char *>(bufs[i0 + k]);
        for (int k = a.n; k < kPackMax; k++) a.buf[k] = nullptr;
        if (pack) {
          a.off_a = h->send_up_off; a.bytes_a = h->send_up_bytes; a.stage_a = su + i0 * h->send_up_bytes;
          a.off_b = h->send_down_off; a.bytes_b = h->send_down_bytes; a.stage_b = sd + i0 * h->send_down_bytes;
        } else {
          a.off_a = h->recv_up_off; a.bytes_a = h->recv_up_bytes; a.stage_a = ru + i0 * h->recv_up_bytes;
          a.off_b = h->recv_down_off; a.bytes_b = h->recv_down_bytes; a.stage_b = rd + i0 * h->recv_down_bytes;
        }
        const size_t big = a.bytes_a > a.bytes_b ? a.bytes_a : a.bytes_b;
        if (big == 0) continue;
        const unsigned gx = (unsigned)((big / 16 + 255) / 256 ? (big / 16 + 255) / 256 : 1);
        dim3 grid(gx > 8 ? 8 : gx, (unsigned)a.n, 2);
        if (pack) k_halo_pack<true><<<grid, 256, 0, s>>>(a);
        else k_halo_pack<false><<<grid, 256, 0, s>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
      }
      return cudaSuccess;
    };
    cudaError_t e = run(true);
    if (e != cudaSuccess) return comm_fail(c, IPB_ERR_CUDA, std::string("halo pack: ") + cudaGetErrorString(e));
    ncclResult_t r = n.GroupStart();
    if (r == ncclSuccess && h->send_up_bytes) r = n.Send(su, nbufs * h->send_up_bytes, ncclUint8, c->rank - 1, c->comm, s);
    if (r == ncclSuccess && h->recv_up_bytes) r = n.Recv(ru, nbufs * h->recv_up_bytes, ncclUint8, c->rank - 1, c->comm, s);
    if (r == ncclSuccess && h->send_down_bytes) r = n.Send(sd, nbufs * h->send_down_bytes, ncclUint8, c->rank + 1, c->comm, s);
    if (r == ncclSuccess && h->recv_down_bytes) r = n.Recv(rd, nbufs * h->recv_down_bytes, ncclUint8, c->rank + 1, c->comm, s);
    ncclResult_t r2 = n.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return comm_fail(c, IPB_ERR_CUDA, std::string("halo exchange: ") + n.GetErrorString(r));
    e = run(false);
    if (e != cudaSuccess) return comm_fail(c, IPB_ERR_CUDA, std::string("halo unpack: ") + cudaGetErrorString(e));
    return IPB_OK;
  }

  ncclResult_t r = n.GroupStart();
  for (size_t i = 0; r == ncclSuccess && i < nbufs; i++) {
    unsigned char *b = static_cast<unsigned char *>(bufs[i]);
    if (h->send_up_bytes) r = n.Send(b + h->send_up_off, h->send_up_bytes, ncclUint8, c->rank - 1, c->comm, s);
    if (r == ncclSuccess && h->recv_up_bytes) r = n.Recv(b + h->recv_up_off, h->recv_up_bytes, ncclUint8, c->rank - 1, c->comm, s);
    if (r == ncclSuccess && h->send_down_bytes) r = n.Send(b + h->send_down_off, h->send_down_bytes, ncclUint8, c->rank + 1, c->comm, s);
    if (r == ncclSuccess && h->recv_down_bytes) r = n.Recv(b + h->recv_down_off, h->recv_down_bytes, ncclUint8, c->rank + 1, c->comm, s);
  }
  ncclResult_t r2 = n.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) return comm_fail(c, IPB_ERR_CUDA, std::string("halo exchange: ") + n.GetErrorString(r));
  return IPB_OK;
}

}  // extern "C"
