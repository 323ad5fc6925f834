// ipb_comm.cu — halo exchange of the row-stripe path (BASELINE config 5) inside the C ABI, on NCCL.
//
// One large frame is cut into row stripes, one per GPU (one process per GPU).  demosaic::full needs one raw row above
// and below a stripe (demosaic.rs:70-74), scaled_demosaic the rows of its windows (scaling.rs:77-87): those rows are
// the only data that travels — grouped ncclSend / ncclRecv between stripe neighbours, enqueued on the context's
// stream (stream-ordered with the kernels; capturable into a CUDA graph).  libnccl is resolved at run time
// (dlopen("libnccl.so.2")): a process that already carries an NCCL — e.g. the one PyTorch ships — shares that copy,
// any other host gets the system's; libipb200.so itself links neither NCCL nor torch.
#include "../../include/ipb200.h"

#include <dlfcn.h>
#include <nccl.h>

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
};

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
