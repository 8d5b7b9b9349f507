// Library-owned gradient exchange: one NCCL communicator per process (one process per GPU), collectives on a private
// communication stream that is forked from / joined to the caller's compute stream with events, so that a bucket's
// all-reduce overlaps the rest of the backward pass — eagerly and inside a captured CUDA graph alike (event record /
// wait pairs between a capturing stream and another stream pull that stream into the capture).
//
// Replaces, for the trainers of this package, what the reference gets from torch.nn.parallel.DistributedDataParallel
// (models/base_model.py:725-737: per-bucket all-reduce overlapped with backward, mean over ranks); the 1/world factor
// is folded into the fused optimizer kernel.  NCCL is bound at run time (dlopen of libnccl.so.2: in a PyTorch process
// that is the copy torch already loaded), so the library itself links against nothing but cudart and still loads on a
// machine without NCCL — jg_comm_* then fail loudly.
#include <dlfcn.h>
#include <mutex>
#include <string.h>

#include "common.cuh"

namespace jg {

// the few NCCL declarations this file needs (nccl.h, stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclSum = 0 };
enum { ncclFloat32 = 7, ncclBfloat16 = 9, ncclUint8 = 1 };

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(int);
  int (*GetVersion)(int*);
  bool ok = false;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.ok) return JG_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  JG_CHECK(h != nullptr, JG_ERR_UNSUPPORTED, "jg_comm: cannot load libnccl.so.2 (%s)", dlerror());
#define JG_SYM(field, name)                                                                   \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                                   \
  JG_CHECK(g_nccl.field != nullptr, JG_ERR_UNSUPPORTED, "jg_comm: libnccl lacks %s", name)
  JG_SYM(GetUniqueId, "ncclGetUniqueId");
  JG_SYM(CommInitRank, "ncclCommInitRank");
  JG_SYM(CommDestroy, "ncclCommDestroy");
  JG_SYM(AllReduce, "ncclAllReduce");
  JG_SYM(Broadcast, "ncclBroadcast");
  JG_SYM(GetErrorString, "ncclGetErrorString");
  JG_SYM(GetVersion, "ncclGetVersion");
#undef JG_SYM
  g_nccl.ok = true;
  return JG_OK;
}

#define JG_NCCL(call)                                                                                  \
  do {                                                                                                 \
    int r__ = (call);                                                                                  \
    if (r__ != ncclSuccess) {                                                                          \
      ::jg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));       \
      return JG_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)

}  // namespace jg

using namespace jg;

struct jg_comm {
  ncclComm_t comm;
  cudaStream_t stream;   // communication stream (highest priority: a collective should not queue behind compute)
  cudaEvent_t fork, join;
  int rank, world, device;
  unsigned long long collectives, bytes;
};

extern "C" int jg_comm_unique_id(void* id128_host) {
  JG_CHECK(id128_host != nullptr, JG_ERR_INVALID, "jg_comm_unique_id: null pointer");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  JG_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128_host, &id, sizeof(id));
  return JG_OK;
}

extern "C" int jg_comm_init(const void* id128_host, int rank, int world, jg_comm_t* out) {
  JG_CHECK(id128_host && out && world >= 1 && rank >= 0 && rank < world, JG_ERR_INVALID, "jg_comm_init: bad arguments");
  int rc = load_nccl();
  if (rc) return rc;
  jg_comm* c = new jg_comm();
  memset(c, 0, sizeof(*c));
  c->rank = rank;
  c->world = world;
  JG_CUDA(cudaGetDevice(&c->device));
  int lo = 0, hi = 0;
  JG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  JG_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
  JG_CUDA(cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming));
  JG_CUDA(cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming));
  ncclUniqueId id;
  memcpy(&id, id128_host, sizeof(id));
  JG_NCCL(g_nccl.CommInitRank(&c->comm, world, id, rank));
  *out = c;
  return JG_OK;
}

// comm stream <- everything the compute stream has been given so far
static int fork_from(jg_comm* c, cudaStream_t compute) {
  JG_CUDA(cudaEventRecord(c->fork, compute));
  JG_CUDA(cudaStreamWaitEvent(c->stream, c->fork, 0));
  return JG_OK;
}

extern "C" int jg_comm_allreduce_async(jg_comm_t c, void* buf, size_t count, int dtype, jg_stream_t compute_stream) {
  JG_CHECK(c && buf && count > 0, JG_ERR_INVALID, "jg_comm_allreduce_async: bad arguments");
  JG_CHECK(dtype == 0 || dtype == 1, JG_ERR_INVALID, "jg_comm_allreduce_async: dtype 0 (fp32) or 1 (bf16)");
  int rc = fork_from(c, static_cast<cudaStream_t>(compute_stream));
  if (rc) return rc;
  JG_NCCL(g_nccl.AllReduce(buf, buf, count, dtype == 0 ? ncclFloat32 : ncclBfloat16, ncclSum, c->comm, c->stream));
  c->collectives += 1;
  c->bytes += count * (dtype == 0 ? 4 : 2);
  return JG_OK;
}

extern "C" int jg_comm_broadcast(jg_comm_t c, void* buf, size_t bytes, int root, jg_stream_t compute_stream) {
  JG_CHECK(c && buf && bytes > 0 && root >= 0 && root < c->world, JG_ERR_INVALID, "jg_comm_broadcast: bad arguments");
  int rc = fork_from(c, static_cast<cudaStream_t>(compute_stream));
  if (rc) return rc;
  JG_NCCL(g_nccl.Broadcast(buf, buf, bytes, ncclUint8, root, c->comm, c->stream));
  return jg_comm_wait(c, compute_stream);
}

extern "C" int jg_comm_wait(jg_comm_t c, jg_stream_t compute_stream) {
  JG_CHECK(c != nullptr, JG_ERR_INVALID, "jg_comm_wait: null communicator");
  JG_CUDA(cudaEventRecord(c->join, c->stream));
  JG_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(compute_stream), c->join, 0));
  return JG_OK;
}

extern "C" int jg_comm_info(jg_comm_t c, int* rank, int* world, unsigned long long* collectives,
                            unsigned long long* bytes, int* nccl_version) {
  JG_CHECK(c != nullptr, JG_ERR_INVALID, "jg_comm_info: null communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (collectives) *collectives = c->collectives;
  if (bytes) *bytes = c->bytes;
  if (nccl_version) g_nccl.GetVersion(nccl_version);
  return JG_OK;
}

extern "C" int jg_comm_destroy(jg_comm_t c) {
  if (!c) return JG_OK;
  cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  cudaEventDestroy(c->fork);
  cudaEventDestroy(c->join);
  cudaStreamDestroy(c->stream);
  delete c;
  return JG_OK;
}
