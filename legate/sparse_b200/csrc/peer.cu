// peer.cu -- NVLink peer-memory exchange kernels (one process per GPU, buffers shared with CUDA IPC).
//
// The SpMV / CG hot path needs two tiny exchanges per step: the x halo (a few KB for stencils) and 1..4
// scalars (CG inner products).  Through NCCL each costs ~15 us of latency; here the ranks write straight into
// each other's memory over NVLink/NVSwitch and synchronise with epoch flags in the same buffers:
//
//   b2s_peer_allreduce : one-shot all-reduce of <= 4 scalars.  Every rank stores its values + an epoch flag
//                        into every peer's mailbox, spins on its own mailbox, then sums in rank order
//                        (identical, deterministic result on every rank).
//   b2s_peer_halo_exchange : copies slices of the local x straight into the neighbours' x buffers (remote
//                        stores), then raises a flag there; a wait kernel spins until all expected
//                        flags of the epoch have arrived.  Acknowledgements protect the next push (WAR).
//   b2s_peer_push      : the same push with many CTAs per destination and NO wait kernel (fused protocol): the
//                        consumers are the SpMV kernels themselves (b2s_spmv_csr_fused polls the arrival flags
//                        right before the first tile that needs them), so the transfer hides behind the tiles
//                        that read local columns.  For small halos the SpMV kernel also does the push itself
//                        (spmv_kernels.cuh), making exchange + product one graph-replayable launch.
//
// Replaces the implicit Legion/Realm copies and future-map reductions of the reference
// (sparse/partition.py:139-208 halo windows; sparse/linalg.py:540,550 dot futures).
// Every rank's IPC buffer starts with a PeerHeader (8 KiB); user data (the x vector) follows at PEER_DATA_OFF.
// Spins are bounded: on timeout the kernel sets PeerHeader::error and returns instead of hanging the GPU.
#include "common.cuh"

namespace b2s {

constexpr int PEER_MAX = 16;
constexpr int PEER_DATA_OFF = 8192;
constexpr long long SPIN_LIMIT = 1LL << 28;  // ~30 s of polling (20 ns sleep + a system-scope load per try): ranks may be
                                             // seconds apart (graph capture, host-side work) without it being an error

struct PeerHeader {
  unsigned long long ar_epoch;                  // completed all-reduces (local, device-incremented)
  unsigned long long halo_epoch;                // completed halo exchanges
  unsigned long long error;                     // != 0 after a spin timeout
  unsigned long long pad0[13];
  unsigned long long ar_flag[2][PEER_MAX];      // written by peers: ar_flag[parity][src] = epoch
  double ar_val[2][PEER_MAX][4];                // written by peers
  unsigned long long halo_flag[PEER_MAX];       // written by peers: epoch of the last halo pushed by src
  unsigned long long halo_ack[PEER_MAX];        // written by peers: epoch dst has finished consuming
  // second, independent protocol: exchanges whose consumer is the SpMV kernel itself (b2s_spmv_csr_fused waits on
  // fuse_flag inside the kernel).  The pusher is either that same kernel (halo mode) or b2s_peer_push (all-gather
  // mode, many CTAs per destination).  Device-side epoch: fuse_epoch = exchanges completed by this rank.
  unsigned long long fuse_flag[PEER_MAX];       // written by peers: epoch of the last slice pushed here by src
  unsigned long long fuse_ack[PEER_MAX];        // written by peers: epoch dst has finished consuming
  unsigned long long fuse_epoch;                // local
  unsigned int fuse_ticket;                     // last-CTA election of the fused SpMV kernel
  unsigned int push_ticket_all;                 // last-CTA election of b2s_peer_push
  unsigned int push_ticket[PEER_MAX];           // per destination: CTAs of b2s_peer_push that finished their part
};
static_assert(sizeof(PeerHeader) <= PEER_DATA_OFF, "header must fit before the data region");

struct PeerPtrs { unsigned char* p[PEER_MAX]; };
struct HaloSends {
  int n;
  int peer[PEER_MAX];
  long long src_off[PEER_MAX], dst_off[PEER_MAX], count[PEER_MAX];  // element offsets inside the x buffers
};
struct HaloRecvs { int n; int peer[PEER_MAX]; };

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until_ge(const unsigned long long* flag, unsigned long long want, PeerHeader* me) {
  long long n = 0;
  while (ld_sys(flag) < want) {
    if (++n > SPIN_LIMIT) { me->error = 1; return false; }
    __nanosleep(20);
  }
  return true;
}

template <typename V>
__global__ void __launch_bounds__(32)
peer_allreduce_kernel(PeerPtrs peers, int rank, int nranks, V* inout, int count) {
  PeerHeader* me = reinterpret_cast<PeerHeader*>(peers.p[rank]);
  const int lane = threadIdx.x;
  const unsigned long long e = me->ar_epoch + 1;
  const int par = (int)(e & 1);
  if (lane < nranks) {
    PeerHeader* dst = reinterpret_cast<PeerHeader*>(peers.p[lane]);
    for (int c = 0; c < count; c++) dst->ar_val[par][rank][c] = (double)inout[c];
    __threadfence_system();
    st_sys(&dst->ar_flag[par][rank], e);
    spin_until_ge(&me->ar_flag[par][lane], e, me);
  }
  __syncwarp();
  if (lane == 0) {
    for (int c = 0; c < count; c++) {
      double s = 0.0;
      for (int q = 0; q < nranks; q++) s += *reinterpret_cast<volatile double*>(&me->ar_val[par][q][c]);
      inout[c] = (V)s;
    }
    me->ar_epoch = e;
  }
}

template <typename V>
__global__ void __launch_bounds__(256)
peer_halo_push_kernel(PeerPtrs peers, int rank, const V* __restrict__ x_local, HaloSends sends, HaloRecvs recvs,
                      long long data_off) {
  PeerHeader* me = reinterpret_cast<PeerHeader*>(peers.p[rank]);
  // epoch: device counter, advanced by the wait kernel
  const unsigned long long e = me->halo_epoch + 1;
  const int b = blockIdx.x;
  // acknowledge epoch e-1 to everyone who pushed to me: this kernel is stream-ordered after the SpMV that
  // consumed those halos, so their buffers may be overwritten now
  if (b == 0 && threadIdx.x < recvs.n) {
    PeerHeader* src = reinterpret_cast<PeerHeader*>(peers.p[recvs.peer[threadIdx.x]]);
    st_sys(&src->halo_ack[rank], e - 1);
  }
  if (b >= sends.n) return;
  const int q = sends.peer[b];
  __shared__ bool ok;
  if (threadIdx.x == 0) ok = spin_until_ge(&me->halo_ack[q], e - 1, me);
  __syncthreads();
  if (!ok) return;
  V* dst = reinterpret_cast<V*>(peers.p[q] + data_off) + sends.dst_off[b];
  const V* src = x_local + sends.src_off[b];
  const long long n = sends.count[b];
  const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
  constexpr int PER = 16 / (int)sizeof(V);
  long long done = 0;
  if (vec) {
    const long long nv = n / PER;
    for (long long i = threadIdx.x; i < nv; i += blockDim.x)
      reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(src)[i];
    done = nv * PER;
  }
  for (long long i = done + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    PeerHeader* dh = reinterpret_cast<PeerHeader*>(peers.p[q]);
    st_sys(&dh->halo_flag[rank], e);
  }
}

// Multi-CTA push of the fused protocol: CTA (s, c) copies chunk c of slice s; the last CTA of a slice raises the
// arrival flag at the destination, the last CTA of the launch advances fuse_epoch.
template <typename V>
__global__ void __launch_bounds__(256)
peer_push_kernel(PeerPtrs peers, int rank, const V* __restrict__ x_local, HaloSends sends, HaloRecvs recvs,
                 long long data_off, int cps) {
  PeerHeader* me = reinterpret_cast<PeerHeader*>(peers.p[rank]);
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&me->fuse_epoch) + 1;
  if (blockIdx.x == 0 && threadIdx.x < recvs.n) {
    // stream-ordered after the kernels that consumed exchange e-1: the senders may overwrite it now
    PeerHeader* src = reinterpret_cast<PeerHeader*>(peers.p[recvs.peer[threadIdx.x]]);
    st_sys(&src->fuse_ack[rank], e - 1);
  }
  __shared__ bool ok;
  __shared__ bool last_of_slice;
  if (sends.n > 0) {
    const int sidx = blockIdx.x / cps, c = blockIdx.x % cps;
    const int q = sends.peer[sidx];
    if (threadIdx.x == 0) ok = spin_until_ge(&me->fuse_ack[q], e - 1, me);
    __syncthreads();
    if (ok) {
      V* dst = reinterpret_cast<V*>(peers.p[q] + data_off) + sends.dst_off[sidx];
      const V* src = x_local + sends.src_off[sidx];
      const long long n = sends.count[sidx];
      constexpr int PER = 16 / (int)sizeof(V);
      // chunk boundaries in whole 16-byte groups of the SOURCE alignment
      const long long per = ((n + cps - 1) / cps + PER - 1) / PER * PER;
      const long long lo = (long long)c * per, hi = lo + per < n ? lo + per : n;
      if (hi > lo) {
        const V* sp = src + lo;
        V* dp = dst + lo;
        const long long m = hi - lo;
        const bool vec = ((reinterpret_cast<uintptr_t>(dp) | reinterpret_cast<uintptr_t>(sp)) & 15) == 0;
        long long done = 0;
        if (vec) {
          const long long nv = m / PER;
          for (long long i = threadIdx.x; i < nv; i += blockDim.x)
            reinterpret_cast<int4*>(dp)[i] = __ldg(reinterpret_cast<const int4*>(sp) + i);
          done = nv * PER;
        }
        for (long long i = done + threadIdx.x; i < m; i += blockDim.x) dp[i] = sp[i];
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int t = atomicAdd(&me->push_ticket[sidx], 1u);
      last_of_slice = (t == (unsigned)cps - 1);
      if (last_of_slice) {
        me->push_ticket[sidx] = 0u;
        __threadfence_system();
        if (ok) {
          PeerHeader* dh = reinterpret_cast<PeerHeader*>(peers.p[q]);
          st_sys(&dh->fuse_flag[rank], e);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&me->push_ticket_all, 1u);
    if (t == gridDim.x - 1) {
      me->push_ticket_all = 0u;
      *reinterpret_cast<volatile unsigned long long*>(&me->fuse_epoch) = e;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(32)
peer_halo_wait_kernel(PeerPtrs peers, int rank, HaloRecvs recvs) {
  PeerHeader* me = reinterpret_cast<PeerHeader*>(peers.p[rank]);
  const unsigned long long e = me->halo_epoch + 1;
  if ((int)threadIdx.x < recvs.n) spin_until_ge(&me->halo_flag[recvs.peer[threadIdx.x]], e, me);
  __syncwarp();
  if (threadIdx.x == 0) { __threadfence_system(); me->halo_epoch = e; }
}

// wait side of b2s_peer_push for consumers that are not b2s_spmv_csr_fused launches: one warp polls the arrival flag of
// every listed source until it reaches the epoch the preceding push advanced to.
__global__ void __launch_bounds__(32)
peer_push_wait_kernel(PeerPtrs peers, int rank, HaloRecvs recvs) {
  PeerHeader* me = reinterpret_cast<PeerHeader*>(peers.p[rank]);
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&me->fuse_epoch);
  if ((int)threadIdx.x < recvs.n) spin_until_ge(&me->fuse_flag[recvs.peer[threadIdx.x]], e, me);
  __syncwarp();
  if (threadIdx.x == 0) __threadfence_system();
}

static int fill_peers(PeerPtrs* pp, int rank, int nranks, void* const* peers_host) {
  B2S_CHECK_ARG(nranks >= 1 && nranks <= PEER_MAX, "nranks %d out of range [1,%d]", nranks, PEER_MAX);
  B2S_CHECK_ARG(rank >= 0 && rank < nranks, "rank %d out of range", rank);
  B2S_CHECK_ARG(peers_host != nullptr, "peer pointer table is NULL");
  for (int i = 0; i < PEER_MAX; i++) pp->p[i] = nullptr;
  for (int i = 0; i < nranks; i++) {
    B2S_CHECK_ARG(peers_host[i] != nullptr, "peer %d pointer is NULL", i);
    pp->p[i] = (unsigned char*)peers_host[i];
  }
  return B2S_OK;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int64_t b2s_peer_header_bytes(void) { return PEER_DATA_OFF; }

int b2s_ipc_alloc(int64_t bytes, void** dev_ptr) {
  B2S_CHECK_ARG(bytes > 0 && dev_ptr != nullptr, "bad size / NULL out pointer");
  void* p = nullptr;
  B2S_CUDA(cudaMalloc(&p, (size_t)bytes));
  B2S_CUDA(cudaMemset(p, 0, (size_t)bytes));
  B2S_CUDA(cudaDeviceSynchronize());
  *dev_ptr = p;
  return B2S_OK;
}

int b2s_ipc_free(void* dev_ptr) {
  if (!dev_ptr) return B2S_OK;
  B2S_CUDA(cudaFree(dev_ptr));
  return B2S_OK;
}

int b2s_peer_allreduce(int vt, int rank, int nranks, void* const* peers_host, void* inout_dev, int count,
                       void* stream) {
  B2S_CHECK_ARG(vt == B2S_F32 || vt == B2S_F64, "bad value type code %d", vt);
  B2S_CHECK_ARG(inout_dev != nullptr && count >= 1 && count <= 4, "count must be 1..4 and inout non-NULL");
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, rank, nranks, peers_host)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (vt == B2S_F32) peer_allreduce_kernel<float><<<1, 32, 0, st>>>(pp, rank, nranks, (float*)inout_dev, count);
  else               peer_allreduce_kernel<double><<<1, 32, 0, st>>>(pp, rank, nranks, (double*)inout_dev, count);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

/* desc: nsends x {peer, src_elem_off, dst_elem_off, count}; recv_peers: ranks that push into this rank.
 * x_local_dev: this rank's x buffer (the data region of its own IPC buffer or any local array). */
static int parse_halo(int rank, int nranks, int nsends, const int64_t* send_desc_host, int nrecvs,
                      const int32_t* recv_peers_host, HaloSends* s, HaloRecvs* r) {
  B2S_CHECK_ARG(nsends >= 0 && nsends <= PEER_MAX && nrecvs >= 0 && nrecvs <= PEER_MAX, "too many halo pieces");
  B2S_CHECK_ARG(nsends == 0 || send_desc_host, "NULL send descriptors");
  B2S_CHECK_ARG(nrecvs == 0 || recv_peers_host, "NULL recv peer list");
  s->n = nsends;
  for (int i = 0; i < nsends; i++) {
    s->peer[i] = (int)send_desc_host[4 * i];
    B2S_CHECK_ARG(s->peer[i] >= 0 && s->peer[i] < nranks && s->peer[i] != rank, "bad destination rank in send %d", i);
    s->src_off[i] = send_desc_host[4 * i + 1];
    s->dst_off[i] = send_desc_host[4 * i + 2];
    s->count[i] = send_desc_host[4 * i + 3];
    B2S_CHECK_ARG(s->count[i] >= 0, "negative count in send %d", i);
  }
  r->n = nrecvs;
  for (int i = 0; i < nrecvs; i++) {
    r->peer[i] = recv_peers_host[i];
    B2S_CHECK_ARG(r->peer[i] >= 0 && r->peer[i] < nranks && r->peer[i] != rank, "bad source rank in recv %d", i);
  }
  return B2S_OK;
}

int b2s_peer_halo_exchange(int vt, int rank, int nranks, void* const* peers_host, const void* x_local_dev,
                           int nsends, const int64_t* send_desc_host, int nrecvs, const int32_t* recv_peers_host,
                           void* stream) {
  B2S_CHECK_ARG(vt == B2S_F32 || vt == B2S_F64, "bad value type code %d", vt);
  B2S_CHECK_ARG(nsends == 0 || x_local_dev, "NULL x");
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, rank, nranks, peers_host)) return rc;
  HaloSends s;
  HaloRecvs r;
  if (int rc = parse_halo(rank, nranks, nsends, send_desc_host, nrecvs, recv_peers_host, &s, &r)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = nsends > 0 ? nsends : 1;
  if (vt == B2S_F32) peer_halo_push_kernel<float><<<blocks, 256, 0, st>>>(pp, rank, (const float*)x_local_dev, s, r, PEER_DATA_OFF);
  else               peer_halo_push_kernel<double><<<blocks, 256, 0, st>>>(pp, rank, (const double*)x_local_dev, s, r, PEER_DATA_OFF);
  B2S_LAUNCH_CHECK();
  peer_halo_wait_kernel<<<1, 32, 0, st>>>(pp, rank, r);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

/* All-gather style push for the fused protocol (PeerHeader::fuse_*): slice i of the local x goes to peer[i]'s x
 * buffer, `ctas_per_send` CTAs per destination so one launch can saturate NVLink (a 5 MB slice to each of 7 peers
 * is 35 MB outbound).  No wait kernel: the consumers are SpMV kernels launched with b2s_spmv_csr_fused, which
 * poll fuse_flag[src] themselves, one column block per source, so the products of a block start as soon as
 * ITS slice has landed.  Epoch e = fuse_epoch + 1; CTA 0 first acknowledges e-1 to every rank that pushes
 * here; the last CTA stores fuse_epoch = e (so a following b2s_spmv_csr_fused uses epoch_add = 0). */
int b2s_peer_push(int vt, int rank, int nranks, void* const* peers_host, const void* x_local_dev, int nsends,
                  const int64_t* send_desc_host, int nrecvs, const int32_t* recv_peers_host, int ctas_per_send,
                  void* stream) {
  B2S_CHECK_ARG(vt == B2S_F32 || vt == B2S_F64, "bad value type code %d", vt);
  B2S_CHECK_ARG(nsends == 0 || x_local_dev, "NULL x");
  B2S_CHECK_ARG(ctas_per_send >= 1 && ctas_per_send <= 64, "ctas_per_send out of range [1,64]");
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, rank, nranks, peers_host)) return rc;
  HaloSends s;
  HaloRecvs r;
  if (int rc = parse_halo(rank, nranks, nsends, send_desc_host, nrecvs, recv_peers_host, &s, &r)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = nsends > 0 ? nsends * ctas_per_send : 1;
  if (vt == B2S_F32) peer_push_kernel<float><<<blocks, 256, 0, st>>>(pp, rank, (const float*)x_local_dev, s, r, PEER_DATA_OFF, ctas_per_send);
  else               peer_push_kernel<double><<<blocks, 256, 0, st>>>(pp, rank, (const double*)x_local_dev, s, r, PEER_DATA_OFF, ctas_per_send);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

/* Wait (on `stream`) until the slices of the listed source ranks pushed with b2s_peer_push have landed -- for
 * consumers that read ALL of x (an unsplit shard of a random matrix: every tile needs every slice, so there is nothing
 * to overlap inside the kernel).  Must be stream-ordered after this rank's own b2s_peer_push of the same exchange. */
int b2s_peer_push_wait(int rank, int nranks, void* const* peers_host, int nrecvs, const int32_t* recv_peers_host,
                       void* stream) {
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, rank, nranks, peers_host)) return rc;
  HaloSends s;
  HaloRecvs r;
  if (int rc = parse_halo(rank, nranks, 0, nullptr, nrecvs, recv_peers_host, &s, &r)) return rc;
  peer_push_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, rank, r);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

/* byte offsets inside the peer header (fused protocol): which = 0 fuse_flag[idx], 1 error word, 2 fuse_ack[idx],
 * 3 fuse_epoch, 4 fuse_ticket */
int64_t b2s_peer_header_offset(int which, int idx) {
  if (which == 0 && idx >= 0 && idx < PEER_MAX) return (int64_t)offsetof(PeerHeader, fuse_flag) + 8 * idx;
  if (which == 1) return (int64_t)offsetof(PeerHeader, error);
  if (which == 2 && idx >= 0 && idx < PEER_MAX) return (int64_t)offsetof(PeerHeader, fuse_ack) + 8 * idx;
  if (which == 3) return (int64_t)offsetof(PeerHeader, fuse_epoch);
  if (which == 4) return (int64_t)offsetof(PeerHeader, fuse_ticket);
  return -1;
}

/* reads PeerHeader::error of this rank's buffer (syncs the stream). */
int b2s_peer_check(void* own_buf_dev, void* stream, int64_t* error_out_host) {
  B2S_CHECK_ARG(own_buf_dev && error_out_host, "NULL pointer");
  unsigned long long e = 0;
  B2S_CUDA(cudaMemcpyAsync(&e, (unsigned char*)own_buf_dev + offsetof(PeerHeader, error), sizeof(e),
                           cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  B2S_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *error_out_host = (int64_t)e;
  return B2S_OK;
}

}  // extern "C"
