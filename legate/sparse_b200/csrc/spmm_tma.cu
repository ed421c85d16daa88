// spmm_tma.cu -- CSR x dense for matrices with column locality: persistent, warp-specialised, TMA-staged X WINDOWS.
//
// The gather kernels of spmm.cu fetch one row of X per NONZERO (k*sv bytes through L1/L2, latency-bound: 0.365 of
// the HBM roofline on the reference's dot_microbenchmark -op spmm shape).  When the rows of a tile of R consecutive
// rows of A touch a narrow window of columns [cmin, cmax] (banded / stencil matrices: R + bandwidth columns), the
// rows cmin..cmax of the row-major X are ONE contiguous block of memory.  Here a producer thread streams, per tile and
// STAGES tiles ahead of the consumers, four cp.async.bulk copies into a shared-memory ring: the tile's column indices,
// its values, its row pointers -- and that block of X.  The consumer warps then form the products entirely out of
// shared memory (one 16-byte LDS per nonzero and lane, no global gathers, no tag lookups) and write Y with coalesced
// 16-byte stores.  L2->SM traffic for X drops from nnz*k*sv to (distinct columns per tile)*k*sv, and the copy of tile
// t+1 overlaps the products of tile t.
//
// A small pre-kernel computes [cmin, cmax] per tile on every call (one pass over `indices`; the library stays
// stateless).  Letting the producer warp find the window of the next tile itself was tried and is slower (k = 32 fp64:
// 795 vs 567 us -- the scan's two dependent trips to memory per tile sit on the producer's critical path).  Tiles whose window or nonzeros do not fit a stage (wide column spans, very long rows) are multiplied by
// the consumers straight from global memory, so the kernel is correct for every matrix; the launcher only selects it
// when most tiles fit (`win_ok` fraction), else the gather kernels of spmm.cu run.
// Replaces SpMMCSR::gpu_variant (reference src/sparse/array/csr/spmm.cu:25-110) on that class of matrices.
#include "common.cuh"
#include <limits.h>
#include <mutex>

namespace b2s {

namespace spmmw {

constexpr int NC = 8;                 // consumer warps
constexpr int THREADS = (NC + 1) * 32;
constexpr int STAGES = 2;
constexpr int R = 64;                 // rows of A per tile
constexpr int NCAP = 1024;            // nonzeros staged per tile (+4 alignment slack)
constexpr int WIN_BYTES = 40 * 1024;  // X window per stage

struct __align__(16) TilePlan {
  long long cmin;                     // first column of the window
  int wrows;                          // rows of X in the window; 0 = tile does not fit (direct path)
  int pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "B2S_MMW_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra B2S_MMW_DONE;\n"
      "bra B2S_MMW_WAIT;\n"
      "B2S_MMW_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// [cmin, cmax] of every tile of R rows; wrows = 0 when the tile cannot be staged
template <typename I, typename P>
__global__ void __launch_bounds__(256)
tile_window_kernel(int64_t nrows, int64_t ntiles, const P* __restrict__ indptr, const I* __restrict__ indices,
                   int win_rows_cap, TilePlan* __restrict__ plan, unsigned long long* __restrict__ fit_count) {
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per tile
  if (t >= ntiles) return;
  const int64_t r0 = t * R, r1 = (r0 + R < nrows) ? r0 + R : nrows;
  const int64_t lo = (int64_t)indptr[r0], hi = (int64_t)indptr[r1];
  long long mn = LLONG_MAX, mx = -1;
  // 8 independent loads per lane and round (a tile is ~R * nnz/row indices: two or three rounds, not twenty dependent ones)
  for (int64_t q = lo + lane; q < hi; q += 32 * 8) {
    I c[8];
#pragma unroll
    for (int u = 0; u < 8; u++) c[u] = (q + 32 * u < hi) ? indices[q + 32 * u] : (I)-1;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long v = (long long)c[u];
      if (v >= 0) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
  }
  if (lane == 0) {
    TilePlan e;
    e.cmin = mx >= 0 ? mn : 0;
    const long long w = mx >= 0 ? mx - mn + 1 : 0;
    const bool fits = (hi - lo) <= NCAP && w <= (long long)win_rows_cap;
    e.wrows = fits ? (int)w : 0;
    e.pad = 0;
    plan[t] = e;
    if (fits || hi == lo) atomicAdd(fit_count, 1ull);
  }
}

template <typename V> struct Pk;      // 16-byte pack of the value type
template <> struct Pk<float> { typedef float4 T; static constexpr int N = 4; };
template <> struct Pk<double> { typedef double2 T; static constexpr int N = 2; };
__device__ __forceinline__ void fma_pack(float4& acc, float a, const float4& x) {
  acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
}
__device__ __forceinline__ void fma_pack(double2& acc, double a, const double2& x) {
  acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y);
}
__device__ __forceinline__ float4 zero_pack(float4*) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ double2 zero_pack(double2*) { return make_double2(0.0, 0.0); }

struct __align__(16) StageMeta {
  long long k0, kb, cmin;
  int r0, nr, wrows, last;
};

template <typename V, typename I, typename P>
struct Layout {
  static constexpr int IDX_B = (NCAP + 4) * (int)sizeof(I);
  static constexpr int VAL_B = (NCAP + 4) * (int)sizeof(V);
  static constexpr int RP_B = ((R + 1 + 3 + 3) / 4 * 4) * (int)sizeof(P);
  static constexpr int STAGE_B = (IDX_B + VAL_B + RP_B + WIN_BYTES + 15) / 16 * 16;
  static constexpr int META_OFF = STAGES * STAGE_B;
  static constexpr int BAR_OFF = META_OFF + STAGES * (int)sizeof(StageMeta);
  static constexpr int TOTAL = BAR_OFF + 2 * STAGES * 8;
};

// NP 16-byte packs per lane, lpr = k / (N * NP) lanes per row (a power of two <= 32).  NP > 1 puts more rows of A on a
// warp: the (column, value) broadcast reads and the address arithmetic of a step are shared by 32 / lpr rows, so both the
// instruction count and the shared-memory wavefronts per nonzero drop (k = 32 fp64: 3.5 -> 2.5 instructions, 3 -> 2.5
// wavefronts per nonzero with NP = 2)
template <typename V, typename I, typename P, int NP>
__global__ void __launch_bounds__(THREADS, 2)
spmm_window_tma_kernel(int64_t nrows, int64_t nnz, int64_t ntiles, int k, int lpr_shift, const P* __restrict__ indptr,
                       const I* __restrict__ indices, const V* __restrict__ vals, const V* __restrict__ X, int64_t ldx,
                       V* __restrict__ Y, int64_t ldy, const TilePlan* __restrict__ plan) {
  using LY = Layout<V, I, P>;
  using PT = typename Pk<V>::T;
  constexpr int N = Pk<V>::N;
  extern __shared__ __align__(128) unsigned char smem[];
  StageMeta* metas = reinterpret_cast<StageMeta*>(smem + LY::META_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + LY::BAR_OFF);
  uint64_t* empty = full + STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NC); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      const int64_t nnz4 = nnz & ~(int64_t)3;
      const int64_t rp4 = (nrows + 1) & ~(int64_t)3;
      int it = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int s = it % STAGES;
        const uint32_t par = (uint32_t)((it / STAGES) & 1);
        mbar_wait(&empty[s], par ^ 1u);
        unsigned char* st = smem + (size_t)s * LY::STAGE_B;
        I* sidx = reinterpret_cast<I*>(st);
        V* sval = reinterpret_cast<V*>(st + LY::IDX_B);
        P* srp = reinterpret_cast<P*>(st + LY::IDX_B + LY::VAL_B);
        V* swin = reinterpret_cast<V*>(st + LY::IDX_B + LY::VAL_B + LY::RP_B);
        const int64_t r0 = t * R, r1 = (r0 + R < nrows) ? r0 + R : nrows;
        const TilePlan tp = plan[t];
        const int64_t k0 = (int64_t)indptr[r0], k1 = (int64_t)indptr[r1];
        StageMeta m;
        m.k0 = k0; m.kb = k0 & ~(int64_t)3; m.cmin = tp.cmin; m.r0 = (int)r0; m.nr = (int)(r1 - r0); m.wrows = tp.wrows;
        m.last = 0;
        uint32_t bytes = 0;
        // row pointers [rb, rend) (16-byte groups inside the array; the few trailing ones by plain loads)
        const int64_t rb = r0 & ~(int64_t)3;
        int64_t rend = (r1 + 1 + 3) & ~(int64_t)3;
        if (rend > rp4) rend = rp4;
        if (rend < rb) rend = rb;
        for (int64_t r = rend; r <= r1; r++) srp[r - rb] = indptr[r];
        const uint32_t nrp = (uint32_t)(rend - rb);
        bytes += nrp * (uint32_t)sizeof(P);
        uint32_t nb = 0;
        if (tp.wrows > 0 && k1 > k0) {
          int64_t kend = (k1 + 3) & ~(int64_t)3;
          if (kend > nnz4) kend = nnz4;
          if (kend < m.kb) kend = m.kb;
          for (int64_t q = kend; q < k1; q++) { sidx[q - m.kb] = indices[q]; sval[q - m.kb] = vals[q]; }
          nb = (uint32_t)(kend - m.kb);
          bytes += nb * (uint32_t)(sizeof(I) + sizeof(V)) + (uint32_t)tp.wrows * (uint32_t)k * (uint32_t)sizeof(V);
        }
        metas[s] = m;
        if (bytes) {
          mbar_arrive_expect_tx(&full[s], bytes);
          if (nrp) bulk_g2s(srp, indptr + rb, nrp * (uint32_t)sizeof(P), &full[s]);
          if (nb) {
            bulk_g2s(sidx, indices + m.kb, nb * (uint32_t)sizeof(I), &full[s]);
            bulk_g2s(sval, vals + m.kb, nb * (uint32_t)sizeof(V), &full[s]);
          }
          if (tp.wrows > 0 && k1 > k0)
            bulk_g2s(swin, X + tp.cmin * ldx, (uint32_t)tp.wrows * (uint32_t)k * (uint32_t)sizeof(V), &full[s]);
        } else {
          mbar_arrive(&full[s]);
        }
        it++;
      }
      // sentinel
      const int s = it % STAGES;
      const uint32_t par = (uint32_t)((it / STAGES) & 1);
      mbar_wait(&empty[s], par ^ 1u);
      StageMeta m;
      m.k0 = m.kb = m.cmin = 0; m.r0 = 0; m.nr = 0; m.wrows = 0; m.last = 1;
      metas[s] = m;
      mbar_arrive(&full[s]);
    }
  } else {
    // ===== consumers: group of lpr lanes per row =====
    const int ctid = tid - 32;
    const int lpr = 1 << lpr_shift;
    const int grp = ctid >> lpr_shift, sub = ctid & (lpr - 1);
    const int ngrp = (NC * 32) >> lpr_shift;
    int it = 0;
    while (true) {
      const int s = it % STAGES;
      const uint32_t par = (uint32_t)((it / STAGES) & 1);
      mbar_wait(&full[s], par);
      const StageMeta m = metas[s];
      if (m.last) break;
      unsigned char* st = smem + (size_t)s * LY::STAGE_B;
      const I* sidx = reinterpret_cast<const I*>(st);
      const V* sval = reinterpret_cast<const V*>(st + LY::IDX_B);
      const P* srp = reinterpret_cast<const P*>(st + LY::IDX_B + LY::VAL_B) + (m.r0 & 3);
      const PT* swin = reinterpret_cast<const PT*>(st + LY::IDX_B + LY::VAL_B + LY::RP_B);
      const int packs_row = k / N;              // == lpr * NP
      for (int j = grp; j < m.nr; j += ngrp) {
        const int64_t ps = (int64_t)srp[j], pe = (int64_t)srp[j + 1];
        PT acc[NP];
#pragma unroll
        for (int u = 0; u < NP; u++) acc[u] = zero_pack((PT*)nullptr);
        if (m.wrows > 0) {
          // (column, value) are read by every lane of the group (a shared-memory broadcast); handing them round by
          // shuffle instead was measured and is slower (k = 32 fp64: 896 vs 673 us): the issue slots, not L1TEX, give out.
          // 16-byte broadcast loads of four pairs behind an alignment head loop: 821 us (rows of 11 are mostly head/tail)
          int q = (int)(ps - m.kb);
          const int qe = (int)(pe - m.kb);
          constexpr int UN = NP >= 4 ? 2 : 4;    // nonzeros in flight per group: UN * NP 16-byte loads
          for (; q + UN <= qe; q += UN) {
            PT x[UN][NP];
#pragma unroll
            for (int e = 0; e < UN; e++) {
              const PT* xr = swin + (int)((long long)sidx[q + e] - m.cmin) * packs_row + sub;
#pragma unroll
              for (int u = 0; u < NP; u++) x[e][u] = xr[u * lpr];
            }
#pragma unroll
            for (int e = 0; e < UN; e++) {
              const V a = sval[q + e];
#pragma unroll
              for (int u = 0; u < NP; u++) fma_pack(acc[u], a, x[e][u]);
            }
          }
          for (; q < qe; q++) {
            const PT* xr = swin + (int)((long long)sidx[q] - m.cmin) * packs_row + sub;
            const V a = sval[q];
#pragma unroll
            for (int u = 0; u < NP; u++) fma_pack(acc[u], a, xr[u * lpr]);
          }
        } else {
          // tile too wide / too long for a stage: straight from global memory
          for (int64_t p = ps; p < pe; p++) {
            const PT* xr = reinterpret_cast<const PT*>(X + (int64_t)indices[p] * ldx) + sub;
            const V a = vals[p];
#pragma unroll
            for (int u = 0; u < NP; u++) fma_pack(acc[u], a, __ldg(xr + u * lpr));
          }
        }
        PT* yr = reinterpret_cast<PT*>(Y + (int64_t)(m.r0 + j) * ldy) + sub;
#pragma unroll
        for (int u = 0; u < NP; u++) yr[u * lpr] = acc[u];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      it++;
    }
  }
}

}  // namespace spmmw

// returns B2S_OK and sets *used = 1 when the window kernel ran, *used = 0 when the caller should use the gather kernels
template <typename V, typename I, typename P>
int spmm_window_try(int64_t nrows, int64_t nnz, int64_t k, const void* indptr, const void* indices, const void* vals,
                    const void* X, int64_t ldx, void* Y, int64_t ldy, cudaStream_t st, int force, int* used) {
  using namespace spmmw;
  *used = 0;
  constexpr int N = Pk<V>::N;
  if (nrows <= 0 || nnz <= 0 || k <= 0) return B2S_OK;
  // eligibility: 16-byte packs, a power-of-two number of them per row, contiguous 16-byte aligned rows
  if (k % N != 0) return B2S_OK;
  const int64_t packs = k / N;
  if (packs > 128 || (packs & (packs - 1)) != 0) return B2S_OK;
  if (!force && packs < 8) return B2S_OK;   // X rows under 128 bytes: the gather kernels are as fast (k = 8: 330 vs 343 us)
  // packs per lane: as many as keep 8 lanes on a row (quarter-warp 16-byte accesses stay conflict-free), at most 4
  int np = 1;
  while (np < 4 && packs / (np * 2) >= 8) np *= 2;
  if (const char* e = getenv("B2S_SPMM_NP")) {
    const int v = atoi(e);
    if ((v == 1 || v == 2 || v == 4) && packs % v == 0) np = v;
  }
  const int64_t lpr = packs / np;
  if (lpr > 32) return B2S_OK;
  if (ldx != k || (ldy % N) != 0 || !aligned16(X) || !aligned16(Y) || !aligned16(indices) || !aligned16(vals) || !aligned16(indptr))
    return B2S_OK;
  int shift = 0;
  while ((1 << shift) < lpr) shift++;
  const int win_rows_cap = (int)(WIN_BYTES / (k * (int64_t)sizeof(V)));
  if (win_rows_cap < R / 2) return B2S_OK;
  const int64_t ntiles = (nrows + R - 1) / R;
  DeviceProps pr;
  if (int rc = get_props(&pr)) return rc;
  TilePlan* plan = nullptr;
  B2S_CUDA(cudaMallocAsync((void**)&plan, sizeof(TilePlan) * (size_t)ntiles + 16, st));
  unsigned long long* fit = reinterpret_cast<unsigned long long*>(plan + ntiles);
  B2S_CUDA(cudaMemsetAsync(fit, 0, 8, st));
  {
    const int64_t nthreads = ntiles * 32;
    tile_window_kernel<I, P><<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(nrows, ntiles, (const P*)indptr,
                                                                                (const I*)indices, win_rows_cap, plan, fit);
    B2S_LAUNCH_CHECK();
  }
  if (!force) {
    // The window kernel pays only when most tiles fit a stage.  The verdict for a structure (one 8-byte read-back, the
    // only synchronisation) is remembered per (arrays, shape), so repeated products with one matrix stay asynchronous.
    struct Verdict { const void *ip, *ix; int64_t nrows, nnz, k; int sv, ok; };
    static Verdict cache[32];
    static int next = 0;
    static std::mutex mu;
    int verdict = -1;
    {
      std::lock_guard<std::mutex> g(mu);
      for (const Verdict& e : cache)
        if (e.ip == indptr && e.ix == indices && e.nrows == nrows && e.nnz == nnz && e.k == k && e.sv == (int)sizeof(V)) verdict = e.ok;
    }
    if (verdict < 0) {
      unsigned long long nfit = 0;
      B2S_CUDA(cudaMemcpyAsync(&nfit, fit, 8, cudaMemcpyDeviceToHost, st));
      B2S_CUDA(cudaStreamSynchronize(st));
      verdict = nfit * 10 >= (unsigned long long)ntiles * 9 ? 1 : 0;
      std::lock_guard<std::mutex> g(mu);
      cache[next] = Verdict{indptr, indices, nrows, nnz, k, (int)sizeof(V), verdict};
      next = (next + 1) % 32;
    }
    if (!verdict) {
      B2S_CUDA(cudaFreeAsync(plan, st));
      return B2S_OK;
    }
  }
  using LY = Layout<V, I, P>;
  int64_t grid = (int64_t)pr.sm_count * 2;
  if (grid > ntiles) grid = ntiles;
  auto launch = [&](auto kern, auto tag) -> int {
    if (int rc = ensure_dyn_smem<decltype(tag)>(kern, LY::TOTAL)) return rc;
    kern<<<(unsigned)grid, THREADS, LY::TOTAL, st>>>(nrows, nnz, ntiles, (int)k, shift, (const P*)indptr, (const I*)indices,
                                                     (const V*)vals, (const V*)X, ldx, (V*)Y, ldy, plan);
    return B2S_OK;
  };
  struct TagW1 {}; struct TagW2 {}; struct TagW4 {};
  int lrc;
  if (np == 4) lrc = launch(spmm_window_tma_kernel<V, I, P, 4>, TagW4{});
  else if (np == 2) lrc = launch(spmm_window_tma_kernel<V, I, P, 2>, TagW2{});
  else lrc = launch(spmm_window_tma_kernel<V, I, P, 1>, TagW1{});
  if (lrc) { cudaFreeAsync(plan, st); return lrc; }
  B2S_LAUNCH_CHECK();
  B2S_CUDA(cudaFreeAsync(plan, st));
  *used = 1;
  return B2S_OK;
}

#define B2S_INST(V, I, P)                                                                                              \
  template int spmm_window_try<V, I, P>(int64_t, int64_t, int64_t, const void*, const void*, const void*, const void*, \
                                        int64_t, void*, int64_t, cudaStream_t, int, int*);
B2S_INST(float, int32_t, int32_t) B2S_INST(float, int32_t, int64_t) B2S_INST(float, int64_t, int32_t) B2S_INST(float, int64_t, int64_t)
B2S_INST(double, int32_t, int32_t) B2S_INST(double, int32_t, int64_t) B2S_INST(double, int64_t, int32_t) B2S_INST(double, int64_t, int64_t)
#undef B2S_INST

}  // namespace b2s
