// comm.cu -- NCCL communicator owned by the library, so a non-Python host can drive the sharded path.
//
// The reference owns NCCL natively (src/sparse/sort/sort.cu:163-322 uses the communicator legate.core hands every
// GPU task; sparse/runtime.py:84-87 pre-initialises it).  Here the host passes a 128-byte ncclUniqueId (created
// on rank 0 with b2s_comm_unique_id and shipped to the other ranks by whatever transport the host has) and gets
// an opaque handle back; the two collectives the row-sharded path needs are
//   b2s_allgather_x       : all-gather of the equally sized x shards before an SpMV     (sparse/csr.py:930-968)
//   b2s_allreduce_scalars : fp64 sum of the CG scalars (rho, p.q, ||r||^2)              (sparse/linalg.py:540,550,561)
// NCCL is bound at run time (dlopen of libnccl.so.2, the SONAME both the system 2.27 and torch's bundled 2.28
// export), so the library carries no link-time dependency and shares the NCCL that is already in the process.
// On one NVSwitch box the default data plane of the Python layer is the peer-memory kernels of peer.cu / the
// exchange fused into the SpMV kernel; this is the portable (multi-node capable) path and the C-level entry.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

namespace b2s {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
  std::lock_guard<std::mutex> g(g_nccl_mu);
  if (g_nccl.handle) return B2S_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return B2S_EUNSUPPORTED; }
#define B2S_SYM(field, name)                                                          \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));            \
  if (!g_nccl.field) { set_error("libnccl has no symbol %s", name); dlclose(h); return B2S_EUNSUPPORTED; }
  B2S_SYM(GetUniqueId, "ncclGetUniqueId")
  B2S_SYM(CommInitRank, "ncclCommInitRank")
  B2S_SYM(CommDestroy, "ncclCommDestroy")
  B2S_SYM(AllGather, "ncclAllGather")
  B2S_SYM(AllReduce, "ncclAllReduce")
  B2S_SYM(GetErrorString, "ncclGetErrorString")
  B2S_SYM(GetVersion, "ncclGetVersion")
#undef B2S_SYM
  g_nccl.handle = h;
  return B2S_OK;
}

#define B2S_NCCL(call)                                                                      \
  do {                                                                                      \
    ncclResult_t r__ = (call);                                                              \
    if (r__ != ncclSuccess) {                                                               \
      set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));   \
      return B2S_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

struct CommHandle {
  uint32_t magic;
  int rank, nranks, device;
  ncclComm_t comm;
};
static constexpr uint32_t kCommMagic = 0xB200C0A1u;

}  // namespace b2s

using namespace b2s;

extern "C" {

/* NCCL version the library bound to (e.g. 22703), 0 if libnccl could not be loaded */
int b2s_comm_nccl_version(void) {
  if (load_nccl() != B2S_OK) return 0;
  int v = 0;
  if (g_nccl.GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

/* rank 0: fill 128 bytes with a fresh ncclUniqueId; ship them to every rank (any transport) */
int b2s_comm_unique_id(void* id128_host_out) {
  B2S_CHECK_ARG(id128_host_out != nullptr, "NULL id buffer");
  if (int rc = load_nccl()) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  B2S_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128_host_out, &id, 128);
  return B2S_OK;
}

/* collective over all ranks: communicator on the CURRENT device of the calling thread */
int b2s_comm_init(int rank, int nranks, const void* nccl_unique_id_128B, void** comm_out) {
  B2S_CHECK_ARG(comm_out != nullptr && nccl_unique_id_128B != nullptr, "NULL pointer");
  B2S_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d / nranks %d", rank, nranks);
  *comm_out = nullptr;
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id_128B, 128);
  CommHandle* h = new CommHandle();
  h->magic = kCommMagic; h->rank = rank; h->nranks = nranks; h->comm = nullptr;
  cudaError_t ce = cudaGetDevice(&h->device);
  if (ce != cudaSuccess) { delete h; set_error("cudaGetDevice -> %s", cudaGetErrorString(ce)); return B2S_ECUDA; }
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    delete h;
    set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, nranks, g_nccl.GetErrorString(r));
    return B2S_ECUDA;
  }
  *comm_out = h;
  return B2S_OK;
}

int b2s_comm_destroy(void* comm) {
  CommHandle* h = (CommHandle*)comm;
  if (!h) return B2S_OK;
  B2S_CHECK_ARG(h->magic == kCommMagic, "not a b2s communicator handle");
  h->magic = 0;
  ncclResult_t r = g_nccl.CommDestroy(h->comm);
  delete h;
  if (r != ncclSuccess) { set_error("ncclCommDestroy -> %s", g_nccl.GetErrorString(r)); return B2S_ECUDA; }
  return B2S_OK;
}

/* x_full[q*n_local .. (q+1)*n_local) = rank q's x_local (equal shards: pad the last one; in place allowed when
 * x_local == x_full + rank*n_local).  Enqueued on `stream`. */
int b2s_allgather_x(void* comm, int vt, const void* x_local, int64_t n_local, void* x_full, void* stream) {
  CommHandle* h = (CommHandle*)comm;
  B2S_CHECK_ARG(h && h->magic == kCommMagic, "bad communicator handle");
  B2S_CHECK_ARG(vt == B2S_F32 || vt == B2S_F64, "bad value type code %d", vt);
  B2S_CHECK_ARG(n_local >= 0 && (n_local == 0 || (x_local && x_full)), "bad shard");
  if (n_local == 0) return B2S_OK;
  B2S_NCCL(g_nccl.AllGather(x_local, x_full, (size_t)n_local, vt == B2S_F32 ? ncclFloat32 : ncclFloat64, h->comm,
                            (cudaStream_t)stream));
  return B2S_OK;
}

/* in-place fp64 sum of `count` device scalars over all ranks */
int b2s_allreduce_scalars(void* comm, void* scalars_dev, int count, void* stream) {
  CommHandle* h = (CommHandle*)comm;
  B2S_CHECK_ARG(h && h->magic == kCommMagic, "bad communicator handle");
  B2S_CHECK_ARG(scalars_dev != nullptr && count >= 1, "bad scalar buffer");
  B2S_NCCL(g_nccl.AllReduce(scalars_dev, scalars_dev, (size_t)count, ncclFloat64, ncclSum, h->comm, (cudaStream_t)stream));
  return B2S_OK;
}

}  // extern "C"
