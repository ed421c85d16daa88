// spmv.cu -- CSR SpMV for sm_100a.
//
// Replaces CSRSpMVRowSplit::gpu_variant (reference src/sparse/array/csr/spmv.cu:24-123, a
// cusparseSpMV call) with hand-written row-blocked ("merge-path tiles rounded to row boundaries")
// kernels:
//
//   plan   : the (rows + nnz) work list is cut into tiles of T merge items; tile t owns the rows whose
//            start position indptr[r] + r falls in [t*T, (t+1)*T).  Every tile therefore has <= T rows and
//            all of its rows except possibly the last fit in one shared-memory chunk of CAP = T + 4
//            nonzeros.  The plan is (ntiles + 1) 16-byte entries {first nnz, first row}.
//   kind 1 (default, "TMA"): persistent CTAs; one producer thread streams each tile's indices / vals /
//            indptr slices into a ring of shared-memory stages with cp.async.bulk (TMA, evict-first L2
//            hint) completing on mbarriers, running STAGES tiles ahead of the consumer warps, which
//            gather x through the read-only path, overwrite vals with vals[k]*x[col[k]] in place and reduce.
//   kind 0 ("LDG"): the same tile processed with 128-bit register loads, one tile per CTA.
//   tail   : a last row longer than the chunk is finished by the whole CTA straight from global memory
//            (block reduction) -- so no cross-CTA carries, no atomics, no fix-up pass.
//   reduce : 2^s lanes per row (s chosen per tile from its row length) reduce the parked products:
//            sequential for short rows (same order as the reference's CPU loop, spmv.cc:36-44),
//            warp-shuffle tree for long rows.  Optional fused epilogue: the CG inner product
//            sum_i w[i]*y[i] (deterministic two-stage grid reduction).
//
// HBM-bound by construction: algorithmic bytes per launch are
//   nnz*(sizeof V + sizeof I) + (nrows+1)*sizeof P + ncols*sizeof V + nrows*sizeof V.
#include "common.cuh"
#include <limits.h>

namespace b2s {

struct __align__(16) PlanEntry {
  long long k;   // first nonzero of the tile's first row
  int row;       // first row of the tile
  int pad;
};

// ---------------------------------------------------------------------------------------------
// Tile configurations.  X(ID, KIND, A, B, C, D)
//   KIND 0 (LDG) : A = THREADS, B = GROUPS (4-nnz groups per thread), C = MINB, D = SCALAR mapping flag
//                  CAP = 4*A*B
//   KIND 1 (TMA) : A = consumer warps, B = 16-byte groups per consumer thread, C = STAGES, D = MINB
//                  CAP = (16/sizeof V) * 32*A * B
// T = CAP - 4 merge items per tile.  Config 0 is the default and is instantiated for every index type;
// the others (tuning sweeps, tools/, b2s_spmv_set_config) only for int32 indices/indptr.
// ---------------------------------------------------------------------------------------------
#define B2S_SPMV_CONFIGS(X) \
  X(0, 1, 4, 4, 2, 6)       \
  X(1, 0, 128, 4, 6, 0)     \
  X(2, 0, 256, 4, 3, 0)     \
  X(3, 0, 128, 2, 8, 1)     \
  X(4, 1, 4, 6, 2, 4)       \
  X(5, 1, 4, 8, 2, 3)       \
  X(6, 1, 4, 3, 2, 8)       \
  X(7, 1, 3, 4, 2, 8)       \
  X(8, 1, 6, 4, 2, 4)       \
  X(9, 1, 4, 4, 3, 4)       \
  X(10, 1, 4, 12, 2, 2)     \
  X(11, 1, 4, 16, 2, 1)     \
  X(12, 1, 8, 8, 2, 1)      \
  X(13, 1, 8, 4, 2, 2)
struct TileCfgRt { int kind, a, b, c, d; };
static const TileCfgRt kCfgs[] = {
#define X(ID, K, A, B, C, D) {K, A, B, C, D},
    B2S_SPMV_CONFIGS(X)
#undef X
};
static constexpr int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);
static int g_cfg = -1;   // -1 = automatic (by value type); >= 0 forced by b2s_spmv_set_config
static int g_waves = 0;  // LDG kind: 0 = one tile per CTA; >0 = grid-stride with waves*SMs*occupancy CTAs
                         // TMA kind: CTAs per SM cap (0 = occupancy)
static constexpr int kDefaultCfgF64 = 0;   // 4 consumer warps x 4 groups, 2 stages, 6 CTAs/SM (CAP 1024)
static constexpr int kDefaultCfgF32 = 6;   // 4 consumer warps x 3 groups, 2 stages, 8 CTAs/SM (CAP 1536)
// scattered matrices (plan statistic > 16 distinct x lines per warp gather) want many gathers in flight per thread:
static constexpr int kScatterCfgF64 = 3;   // LDG tiles, 128 threads x 8 nnz, scalar mapping (R32 fp64: 2.01 ms vs 2.09 row-group)
static constexpr int kScatterCfgF32 = 5;   // TMA tiles, 4 warps x 8 groups = 32 gathers/thread (R32 fp32: 1.33 ms vs 1.78)
static inline int resolve_cfg(int vt, bool scattered = false) {
  if (g_cfg >= 0) return g_cfg;
  if (scattered) return vt == B2S_F32 ? kScatterCfgF32 : kScatterCfgF64;
  return vt == B2S_F32 ? kDefaultCfgF32 : kDefaultCfgF64;
}

static inline int cfg_cap(int c, int vt) {
  const TileCfgRt& k = kCfgs[c];
  if (k.kind == 0) return 4 * k.a * k.b;
  return (vt == B2S_F32 ? 4 : 2) * 32 * k.a * k.b;
}
static inline int cfg_T(int c, int vt) { return cfg_cap(c, vt) - 4; }

// ---------------------------------------------------------------------------------------------
// Plan kernel: plan[t] = {indptr[r], r} for the first row r with indptr[r] + r >= t*T;
// plan[ntiles] = {nnz, nrows}.
// ---------------------------------------------------------------------------------------------
template <typename P>
__global__ void spmv_plan_kernel(int64_t nrows, const P* __restrict__ indptr, int64_t T, int64_t ntiles,
                                 PlanEntry* __restrict__ plan) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  int64_t lo = 0, hi = nrows;  // answer in [0, nrows]
  if (t == ntiles) {
    lo = nrows;
  } else {
    const int64_t target = t * T;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((int64_t)indptr[mid] + mid >= target) hi = mid; else lo = mid + 1;
    }
  }
  PlanEntry e;
  e.k = (long long)indptr[lo];
  e.row = (int)lo;
  e.pad = 0;
  plan[t] = e;
}

// Second plan pass: plan[t].pad = L when every row of tile t has the same length L > 0 (ELL-like tiles:
// fixed-degree graphs, interior rows of banded matrices), else 0.  Such tiles can be reduced straight from
// registers (see the uniform fast path of spmv_tma_kernel).
template <typename P>
__global__ void spmv_plan_uniform_kernel(const P* __restrict__ indptr, int64_t ntiles, PlanEntry* __restrict__ plan,
                                         int ept, unsigned long long* __restrict__ qualifying) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int r0 = plan[t].row, r1 = plan[t + 1].row;
  int L = 0;
  if (r1 > r0) {
    const long long len0 = (long long)indptr[r0 + 1] - (long long)indptr[r0];
    bool same = len0 > 0 && len0 < 32768;
    for (int r = r0 + 1; same && r < r1; r++) same = ((long long)indptr[r + 1] - (long long)indptr[r]) == len0;
    L = same ? (int)len0 : 0;
  }
  plan[t].pad = L;
  // tiles the register fast path can take: L = ept * 2^s (s <= 5) and a 16-byte aligned first nonzero
  const int lpr = L / ept;
  if (L > 0 && lpr * ept == L && lpr <= 32 && (lpr & (lpr - 1)) == 0 && (plan[t].k & 3) == 0)
    atomicAdd(qualifying, 1ull);
}

__device__ __forceinline__ PlanEntry ld_plan(const PlanEntry* p) {
  int4 v = __ldg(reinterpret_cast<const int4*>(p));
  PlanEntry e;
  e.k = ((long long)(unsigned)v.x) | ((long long)v.y << 32);
  e.row = v.z;
  e.pad = v.w;
  return e;
}

// lanes-per-row rule shared by both kernels
__device__ __forceinline__ int lanes_per_row_shift(int64_t nnz_t, int nr) {
  // lanes per row g = 2^gshift, uniform over the tile: the largest power of two <= L/6 (L = mean row
  // length), i.e. every lane adds ~6..12 parked products sequentially before the shuffle tree.  Short rows
  // (L < 12) get g = 1: a plain sequential walk in the reference's accumulation order (spmv.cc:36-44).
  // (One lane per element -- g = L -- is instruction-bound: ~40 instructions per 32 nonzeros.)
  int gshift = 0;
  const int L = (int)((nnz_t + nr - 1) / nr);
  while (gshift < 5 && (L >> (gshift + 1)) >= 6) gshift++;
  return gshift;
}

// Sum of pr[s + lig + g*t], t = 0..R-1, for one lane of a row group.  With `skew` the walk starts at a
// row-dependent offset and wraps around, so that the lanes of a warp -- which sit a whole row length apart
// in shared memory -- hit different banks even when the row length is a multiple of the bank count
// (32-long fp32 rows would otherwise be an 8-way conflict on every read).  Without it the walk is plain
// left-to-right, the reference's accumulation order.
template <typename V>
__device__ __forceinline__ V row_partial(const V* __restrict__ pr, int s, int e, int lig, int gshift, int j, bool skew) {
  V sum = 0;
  if (!skew) {
    for (int k = s + lig; k < e; k += (1 << gshift)) sum += pr[k];
    return sum;
  }
  const int R = (e - s + (1 << gshift) - 1) >> gshift;  // trips of this lane group
  if (R <= 0) return sum;
  int tt = j & ((1 << (31 - __clz(R))) - 1);             // start offset < R (power-of-two mask: no division)
  for (int t = 0; t < R; t++) {
    const int k = s + lig + (tt << gshift);
    if (k < e) sum += pr[k];
    tt = (tt + 1 == R) ? 0 : tt + 1;
  }
  return sum;
}

// ---------------------------------------------------------------------------------------------
// KIND 0: LDG tile kernel.
// ---------------------------------------------------------------------------------------------
template <typename V, typename I, typename P, int THREADS, int GROUPS, int MINB, bool SCALAR, bool DOT>
__global__ void __launch_bounds__(THREADS, MINB)
spmv_tile_kernel(int64_t ntiles, const P* __restrict__ indptr, const I* __restrict__ indices,
                 const V* __restrict__ vals, const V* __restrict__ x, V* __restrict__ y,
                 const PlanEntry* __restrict__ plan, int vec_ok, const V* __restrict__ w, V* dot_out, void* ws) {
  constexpr int CAP = 4 * THREADS * GROUPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* prod = reinterpret_cast<V*>(smem_raw);                                  // CAP values
  uint16_t* sptr = reinterpret_cast<uint16_t*>(smem_raw + sizeof(V) * CAP);  // <= CAP-3 row offsets (rel. to k0)
  __shared__ double red[32];
  __shared__ V s_tail;
  __shared__ bool s_flag;

  const int tid = threadIdx.x;
  double dot_acc = 0.0;

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const PlanEntry e0 = ld_plan(plan + t), e1 = ld_plan(plan + t + 1);
    const int r0 = e0.row, r1 = e1.row;
    const int nr = r1 - r0;
    if (nr <= 0) continue;  // block-uniform: tile lies inside a long row owned by an earlier tile
    const int64_t k0 = e0.k;
    const int64_t k1 = e1.k;
    const int64_t kb = k0 & ~(int64_t)3;           // 16-byte aligned chunk base
    const int off = (int)(k0 - kb);
    const int64_t kce = (k1 < kb + CAP) ? k1 : kb + CAP;  // end of the staged chunk
    const bool has_tail = k1 > kce;

    __syncthreads();  // previous tile's reduce phase is done with prod/sptr

    // row offsets of this tile, relative to k0, clamped to the chunk (uint16: CAP <= 32768)
    for (int j = tid; j <= nr; j += THREADS) {
      int64_t rel = (int64_t)indptr[r0 + j] - k0;
      int64_t lim = kce - k0;
      sptr[j] = (uint16_t)(rel < lim ? rel : lim);
    }

    // ---- phase A: stream nnz [k0, kce) -> prod[k - kb] ---------------------------------------
    if (SCALAR) {
      constexpr int ITEMS = 4 * GROUPS;
      I c[ITEMS];
      V a[ITEMS];
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const int64_t k = kb + tid + THREADS * j;
        const bool in = (k >= k0) && (k < kce);
        c[j] = in ? ld_stream(indices + k) : (I)0;
        a[j] = in ? ld_stream(vals + k) : (V)0;
      }
      V xv[ITEMS];
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const int64_t k = kb + tid + THREADS * j;
        const bool in = (k >= k0) && (k < kce);
        xv[j] = in ? __ldg(x + c[j]) : (V)0;
      }
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const int e = tid + THREADS * j;
        if (kb + e < kce) prod[e] = a[j] * xv[j];
      }
    } else if (vec_ok) {
      I c[GROUPS][4];
      V a[GROUPS][4];
#pragma unroll
      for (int g = 0; g < GROUPS; g++) {
        const int64_t e = kb + 4 * (int64_t)(tid + THREADS * g);
        if (e >= k0 && e + 4 <= kce) {
          ld_stream4(indices + e, c[g]);
          ld_stream4(vals + e, a[g]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const bool in = (e + q >= k0) && (e + q < kce);
            c[g][q] = in ? ld_stream(indices + e + q) : (I)0;
            a[g][q] = in ? ld_stream(vals + e + q) : (V)0;
          }
        }
      }
      V xv[GROUPS][4];
#pragma unroll
      for (int g = 0; g < GROUPS; g++) {
        const int64_t e = kb + 4 * (int64_t)(tid + THREADS * g);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const bool in = (e + q >= k0) && (e + q < kce);
          xv[g][q] = in ? __ldg(x + c[g][q]) : (V)0;
        }
      }
#pragma unroll
      for (int g = 0; g < GROUPS; g++) {
        const int e = 4 * (tid + THREADS * g);
        if (kb + e < kce) {
          if constexpr (sizeof(V) == 8) {
            // 32 B per lane = two 16-byte stores.  Lanes 4..7 of each quarter-warp write their upper half
            // first, so the eight lanes of one store wavefront cover eight distinct 16-byte bank groups
            // (plain lane order would be a 2-way conflict: lane stride 32 B).
            const double p0 = a[g][0] * xv[g][0], p1 = a[g][1] * xv[g][1];
            const double p2 = a[g][2] * xv[g][2], p3 = a[g][3] * xv[g][3];
            const bool h = (tid >> 2) & 1;
            const double2 lo = make_double2(p0, p1), hi = make_double2(p2, p3);
            double2* dst = reinterpret_cast<double2*>(prod + e);
            dst[h ? 1 : 0] = h ? hi : lo;
            dst[h ? 0 : 1] = h ? lo : hi;
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) prod[e + q] = a[g][q] * xv[g][q];
          }
        }
      }
    } else {
      // unaligned base pointers: scalar coalesced loads
#pragma unroll 4
      for (int e = tid; e < CAP; e += THREADS) {
        const int64_t k = kb + e;
        if (k >= k0 && k < kce) prod[e] = ld_stream(vals + k) * __ldg(x + ld_stream(indices + k));
      }
    }

    // ---- tail: remainder of an over-long last row, straight from global -------------------------
    if (has_tail) {
      V ts0 = 0, ts1 = 0, ts2 = 0, ts3 = 0;
      int64_t k = kce + tid;
      for (; k + 3 * THREADS < k1; k += 4 * THREADS) {
        I c0 = ld_stream(indices + k), c1 = ld_stream(indices + k + THREADS);
        I c2 = ld_stream(indices + k + 2 * THREADS), c3 = ld_stream(indices + k + 3 * THREADS);
        V a0 = ld_stream(vals + k), a1 = ld_stream(vals + k + THREADS);
        V a2 = ld_stream(vals + k + 2 * THREADS), a3 = ld_stream(vals + k + 3 * THREADS);
        ts0 += a0 * __ldg(x + c0); ts1 += a1 * __ldg(x + c1);
        ts2 += a2 * __ldg(x + c2); ts3 += a3 * __ldg(x + c3);
      }
      for (; k < k1; k += THREADS) ts0 += ld_stream(vals + k) * __ldg(x + ld_stream(indices + k));
      double tot = block_sum<THREADS>((double)((ts0 + ts1) + (ts2 + ts3)), red);
      if (tid == 0) s_tail = (V)tot;
    }
    __syncthreads();

    // ---- reduce: per-row sums of the parked products ----------------------------------------------
    const int gshift = lanes_per_row_shift(k1 - k0, nr);
    const bool skew = gshift > 0 || ((((k1 - k0) + nr - 1) / nr) & 1) == 0;  // even / long rows: rotate the walk
    const int g = 1 << gshift;
    const int lig = tid & (g - 1);
    const int grp = tid >> gshift;
    const int ngrp = THREADS >> gshift;
    const V* pr = prod + off;
    for (int base = 0; base < nr; base += ngrp) {
      const int j = base + grp;
      const bool active = j < nr;
      const int s = active ? (int)sptr[j] : 0;
      const int e = active ? (int)sptr[j + 1] : 0;
      V sum = row_partial<V>(pr, s, e, lig, gshift, j, skew);
      for (int o = g >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (active && lig == 0) {
        if (has_tail && j == nr - 1) sum += s_tail;
        y[r0 + j] = sum;
        if (DOT) dot_acc += (double)sum * (double)w[r0 + j];
      }
    }
  }

  if (DOT) {
    double part = block_sum<THREADS>(dot_acc, red);
    if (grid_reduce_is_last<THREADS>(ws, part, red, &s_flag)) {
      double total = grid_reduce_final<THREADS>(ws, red);
      if (tid == 0) *dot_out = (V)total;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// KIND 1: TMA-staged persistent kernel (cp.async.bulk + mbarrier ring, warp-specialised).
// ---------------------------------------------------------------------------------------------
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
      "B2S_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra B2S_DONE;\n"
      "bra B2S_WAIT;\n"
      "B2S_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Tile processing order + optional in-kernel wait for a halo that other GPUs are still pushing into x
// (csrc/peer.cu).  Tiles are visited range by range; ranges [0, n_free) only read locally valid x, the rest
// may read halo columns, so before its first such tile each CTA's producer polls the arrival flags (local
// memory, written remotely by the neighbours' push kernels).  The exchange latency hides behind the interior
// tiles -- compute and collective in ONE kernel.  n_flags == 0: ordinary launch.
struct TileOrder {
  int nranges, n_free, n_flags, pad;
  long long lo[6], hi[6];
  const unsigned long long* flag[8];
  unsigned long long expect;
  unsigned long long* error;
};

struct __align__(16) TileMeta {
  long long k0, k1, kb;
  int r0, nr, rb, pad;
};

template <typename V, typename I, typename P, int NC, int G, int STAGES>
struct TmaLayout {
  static constexpr int CT = NC * 32;
  static constexpr int EPT = 16 / (int)sizeof(V);
  static constexpr int CAP = EPT * CT * G;
  static constexpr int T = CAP - 4;
  static constexpr int RPN = ((T + 1 + 3 + 3) / 4) * 4;  // row-pointer slice capacity (incl. alignment slack)
  static constexpr int COLS_B = CAP * (int)sizeof(I);
  static constexpr int VALS_B = CAP * (int)sizeof(V);
  static constexpr int RP_B = RPN * (int)sizeof(P);
  static constexpr int STAGE_B = COLS_B + VALS_B + RP_B;
  static constexpr int META_OFF = STAGES * STAGE_B;
  static constexpr int BAR_OFF = META_OFF + STAGES * (int)sizeof(TileMeta);
  static constexpr int TOTAL = BAR_OFF + 2 * STAGES * 8;
};

template <typename V, typename I, typename P, int NC, int G, int STAGES, int MINB, bool UNI, bool DOT>
__global__ void __launch_bounds__((NC + 1) * 32, MINB)
spmv_tma_kernel(TileOrder order, int64_t nrows, int64_t nnz, const P* __restrict__ indptr,
                const I* __restrict__ indices, const V* __restrict__ vals, const V* __restrict__ x,
                V* __restrict__ y, const PlanEntry* __restrict__ plan, const V* __restrict__ w, V* dot_out, void* ws) {
  using LY = TmaLayout<V, I, P, NC, G, STAGES>;
  constexpr int CT = LY::CT, EPT = LY::EPT, CAP = LY::CAP;
  constexpr int THREADS = (NC + 1) * 32;
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn;
  TileMeta* metas = reinterpret_cast<TileMeta*>(smem_raw + LY::META_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + LY::BAR_OFF);
  uint64_t* empty = full + STAGES;
  __shared__ double red[32];
  __shared__ double cred[NC];
  __shared__ bool s_flag;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  double dot_acc = 0.0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NC);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ===== producer: lane 0 drives the TMA engine, STAGES tiles ahead of the consumers; the other lanes
    // walk the same loop so the warp stays converged for the block-wide barrier of the DOT epilogue =====
    const uint64_t pol = l2_evict_first_policy();
    const int64_t nnz4 = nnz & ~(int64_t)3;
    const int64_t np1 = nrows + 1;
    const int64_t rp4 = np1 & ~(int64_t)3;
    int it = 0;
    long long total = 0, free_total = 0;
    for (int r = 0; r < order.nranges; r++) {
      total += order.hi[r] - order.lo[r];
      if (r < order.n_free) free_total = total;
    }
    bool halo_ready = order.n_flags == 0;
    for (long long v = blockIdx.x; v < total; v += gridDim.x) {
      long long off = v;
      int r = 0;
      while (off >= order.hi[r] - order.lo[r]) { off -= order.hi[r] - order.lo[r]; r++; }
      const int64_t t = order.lo[r] + off;
      if (!halo_ready && v >= free_total) {
        // first tile of this CTA that may read halo columns: wait (bounded) for every neighbour's push
        if (lane == 0) {
          for (int f = 0; f < order.n_flags; f++) {
            long long spins = 0;
            unsigned long long seen;
            do {
              asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(order.flag[f]) : "memory");
              if (seen >= order.expect) break;
              __nanosleep(20);
            } while (++spins < (1LL << 28));
            if (seen < order.expect) *order.error = 1ull;
          }
        }
        __syncwarp();
        halo_ready = true;
      }
      const PlanEntry e0 = ld_plan(plan + t), e1 = ld_plan(plan + t + 1);
      const int nr = e1.row - e0.row;
      if (nr <= 0) continue;  // warp-uniform
      if (lane == 0) {
        const int s = it % STAGES;
        const uint32_t par = (uint32_t)((it / STAGES) & 1);
        mbar_wait(&empty[s], par ^ 1u);
        unsigned char* st = smem_raw + (size_t)s * LY::STAGE_B;
        I* scols = reinterpret_cast<I*>(st);
        V* svals = reinterpret_cast<V*>(st + LY::COLS_B);
        P* srp = reinterpret_cast<P*>(st + LY::COLS_B + LY::VALS_B);
        const int64_t k0 = e0.k, k1 = e1.k;
        const int64_t kb = k0 & ~(int64_t)3;
        const int64_t kce = (k1 < kb + CAP) ? k1 : kb + CAP;
        int64_t kend = (kce + 3) & ~(int64_t)3;  // bulk range [kb, kend): whole 16-byte groups inside the array
        if (kend > nnz4) kend = nnz4;
        if (kend < kb) kend = kb;
        const int64_t r0 = e0.row, r1 = e1.row;
        const int64_t rb = r0 & ~(int64_t)3;
        int64_t rend = (r1 + 1 + 3) & ~(int64_t)3;  // row pointers [rb, rend) by bulk copy
        if (rend > rp4) rend = rp4;
        if (rend < rb) rend = rb;
        TileMeta m;
        m.k0 = k0; m.k1 = k1; m.kb = kb; m.r0 = (int)r0; m.nr = nr; m.rb = (int)rb; m.pad = e0.pad;  // pad = uniform row length
        metas[s] = m;
        // the (at most 3) trailing elements that do not fill a 16-byte group at the very end of an array
        for (int64_t k = kend; k < kce; k++) { scols[k - kb] = indices[k]; svals[k - kb] = vals[k]; }
        for (int64_t r = rend; r <= r1; r++) srp[r - rb] = indptr[r];
        const uint32_t nb = (uint32_t)(kend - kb);
        const uint32_t nrp = (uint32_t)(rend - rb);
        const uint32_t bytes = nb * (uint32_t)(sizeof(I) + sizeof(V)) + nrp * (uint32_t)sizeof(P);
        if (bytes) {
          mbar_arrive_expect_tx(&full[s], bytes);
          if (nb) {
            bulk_g2s(scols, indices + kb, nb * (uint32_t)sizeof(I), &full[s], pol);
            bulk_g2s(svals, vals + kb, nb * (uint32_t)sizeof(V), &full[s], pol);
          }
          if (nrp) bulk_g2s(srp, indptr + rb, nrp * (uint32_t)sizeof(P), &full[s], pol);
        } else {
          mbar_arrive(&full[s]);
        }
      }
      __syncwarp();
      it++;
    }
    if (lane == 0) {
      // sentinel: tells the consumers there is no more work
      const int s = it % STAGES;
      const uint32_t par = (uint32_t)((it / STAGES) & 1);
      mbar_wait(&empty[s], par ^ 1u);
      TileMeta m;
      m.k0 = m.k1 = m.kb = 0; m.r0 = 0; m.nr = -1; m.rb = 0; m.pad = 0;
      metas[s] = m;
      mbar_arrive(&full[s]);
    }
    __syncwarp();
  } else {
    // ===== consumers =====
    const int ctid = tid - 32;
    const int cwarp = warp - 1;
    int it = 0;
    while (true) {
      const int s = it % STAGES;
      const uint32_t par = (uint32_t)((it / STAGES) & 1);
      mbar_wait(&full[s], par);
      const TileMeta m = metas[s];
      if (m.nr < 0) break;
      unsigned char* st = smem_raw + (size_t)s * LY::STAGE_B;
      const I* scols = reinterpret_cast<const I*>(st);
      V* svals = reinterpret_cast<V*>(st + LY::COLS_B);
      const P* srp = reinterpret_cast<const P*>(st + LY::COLS_B + LY::VALS_B) + (m.r0 - m.rb);
      const int64_t k0 = m.k0, k1 = m.k1, kb = m.kb;
      const int nr = m.nr, r0 = m.r0;
      const int off = (int)(k0 - kb);
      const int64_t kce = (k1 < kb + CAP) ? k1 : kb + CAP;
      const bool has_tail = k1 > kce;
      const int lo = off, hi = (int)(kce - kb);  // valid slots [lo, hi)

      // ---- uniform fast path: every row of the tile has the same length L = EPT * 2^s (s <= 5) and the
      // tile starts on a 16-byte group boundary, so each lane's group lies inside one row and a row is a run
      // of 2^s consecutive lanes: sum in registers, shuffle-reduce, store y.  No shared-memory round trip,
      // no row-pointer reads, ~4x fewer instructions per nonzero than the generic reduce below.
      const int UL = m.pad;
      const int lpr = UL / EPT;
      if (UNI && UL > 0 && off == 0 && !has_tail && lpr * EPT == UL && lpr <= 32 && (lpr & (lpr - 1)) == 0) {
        const int lshift = 31 - __clz(lpr);
        I c[G][EPT];
        V a[G][EPT];
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int e = EPT * (ctid + CT * g);
#pragma unroll
          for (int q = 0; q < EPT; q++) { c[g][q] = scols[e + q]; a[g][q] = svals[e + q]; }
        }
        V part[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int e = EPT * (ctid + CT * g);
          V acc = (V)0;
#pragma unroll
          for (int q = 0; q < EPT; q++) acc += (e < hi) ? a[g][q] * __ldg(x + c[g][q]) : (V)0;
          part[g] = acc;
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
          V sum = part[g];
          for (int o = lpr >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const int gi = ctid + CT * g;
          if ((gi & (lpr - 1)) == 0 && EPT * gi < hi) {
            const int row = r0 + (gi >> lshift);
            y[row] = sum;
            if (DOT) dot_acc += (double)sum * (double)w[row];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        it++;
        continue;
      }

      // ---- products in place: svals[e] *= x[scols[e]] ------------------------------------------------
      {
        I c[G][EPT];
        V a[G][EPT];
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int e = EPT * (ctid + CT * g);
#pragma unroll
          for (int q = 0; q < EPT; q++) { c[g][q] = scols[e + q]; a[g][q] = svals[e + q]; }
        }
        V xv[G][EPT];
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int e = EPT * (ctid + CT * g);
#pragma unroll
          for (int q = 0; q < EPT; q++) {
            const bool in = (e + q >= lo) && (e + q < hi);
            xv[g][q] = in ? __ldg(x + c[g][q]) : (V)0;
          }
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int e = EPT * (ctid + CT * g);
          if (e < hi) {
#pragma unroll
            for (int q = 0; q < EPT; q++) svals[e + q] = a[g][q] * xv[g][q];
          }
        }
      }

      // ---- tail of an over-long last row, straight from global ------------------------------------------
      V tail_sum = (V)0;
      if (has_tail) {
        V ts0 = 0, ts1 = 0;
        int64_t k = kce + ctid;
        for (; k + CT < k1; k += 2 * CT) {
          I c0 = ld_stream(indices + k), c1 = ld_stream(indices + k + CT);
          V a0 = ld_stream(vals + k), a1 = ld_stream(vals + k + CT);
          ts0 += a0 * __ldg(x + c0);
          ts1 += a1 * __ldg(x + c1);
        }
        for (; k < k1; k += CT) ts0 += ld_stream(vals + k) * __ldg(x + ld_stream(indices + k));
        double v = warp_sum((double)(ts0 + ts1));
        named_bar_sync(2, CT);  // cred free
        if (lane == 0) cred[cwarp] = v;
        named_bar_sync(2, CT);
        double tot = 0.0;
#pragma unroll
        for (int q = 0; q < NC; q++) tot += cred[q];
        tail_sum = (V)tot;
      }
      named_bar_sync(1, CT);  // all products of this stage are parked

      // ---- reduce ----------------------------------------------------------------------------------------
      const int gshift = lanes_per_row_shift(k1 - k0, nr);
      // even / long rows: rotate the walk (only compiled into the variant the plan selects for such matrices)
      const bool skew = UNI && (gshift > 0 || ((((k1 - k0) + nr - 1) / nr) & 1) == 0);
      const int g = 1 << gshift;
      const int lig = ctid & (g - 1);
      const int grp = ctid >> gshift;
      const int ngrp = CT >> gshift;
      const V* pr = svals + off;
      const int64_t lim = kce - k0;
      for (int base = 0; base < nr; base += ngrp) {
        const int j = base + grp;
        const bool active = j < nr;
        int sidx = 0, eidx = 0;
        if (active) {
          const int64_t a0 = (int64_t)srp[j] - k0, a1 = (int64_t)srp[j + 1] - k0;
          sidx = (int)(a0 < lim ? a0 : lim);
          eidx = (int)(a1 < lim ? a1 : lim);
        }
        V sum = row_partial<V>(pr, sidx, eidx, lig, gshift, j, skew);
        for (int o = g >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (active && lig == 0) {
          if (has_tail && j == nr - 1) sum += tail_sum;
          y[r0 + j] = sum;
          if (DOT) dot_acc += (double)sum * (double)w[r0 + j];
        }
      }
      // release the stage: generic-proxy writes (products) must be ordered before the TMA refills it
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      it++;
    }
  }

  if (DOT) {
    double part = block_sum<THREADS>(dot_acc, red);
    if (grid_reduce_is_last<THREADS>(ws, part, red, &s_flag)) {
      double total = grid_reduce_final<THREADS>(ws, red);
      if (tid == 0) *dot_out = (V)total;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Plan-free fallback: 2^s lanes per row, grid-stride over rows (classic CSR-vector).
// ---------------------------------------------------------------------------------------------
template <typename V, typename I, typename P>
__global__ void __launch_bounds__(256)
spmv_rowgroup_kernel(int64_t nrows, const P* __restrict__ indptr, const I* __restrict__ indices,
                     const V* __restrict__ vals, const V* __restrict__ x, V* __restrict__ y, int gshift) {
  const int g = 1 << gshift;
  const int lig = threadIdx.x & (g - 1);
  const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> gshift;
  const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) >> gshift;
  // all lanes of a warp iterate the same number of times so the shuffles stay converged
  const int64_t warp_first = (((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> gshift);
  for (int64_t rb = warp_first; rb < nrows; rb += ngrp) {
    const int64_t r = rb + (grp - warp_first);
    const bool active = r < nrows;
    const int64_t s = active ? (int64_t)indptr[r] : 0;
    const int64_t e = active ? (int64_t)indptr[r + 1] : 0;
    V sum = 0;
    for (int64_t k = s + lig; k < e; k += g) sum += ld_stream(vals + k) * __ldg(x + ld_stream(indices + k));
    for (int o = g >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (active && lig == 0) y[r] = sum;
  }
}

// ---------------------------------------------------------------------------------------------
// Host-side launchers
// ---------------------------------------------------------------------------------------------
struct SpmvArgs {
  int64_t ntiles, nrows, nnz;
  int64_t tile_lo, tile_hi;  // tile sub-range to run (TMA kernels); [0, ntiles) for a whole SpMV
  const TileOrder* order;    // optional explicit tile order + halo wait (TMA kernels); NULL = [tile_lo, tile_hi)
  const void *indptr, *indices, *vals, *x;
  void* y;
  const PlanEntry* plan;
  int vec_ok;
  int uniform;  // plan found >= 25% ELL-like tiles: use the kernel variant with the register fast path
  const void* w;
  void* dot_out;
  void* ws;
  cudaStream_t st;
};

template <typename V, typename I, typename P, int THREADS, int GROUPS, int MINB, bool SCALAR, bool DOT>
static int launch_ldg(const SpmvArgs& a) {
  constexpr int CAP = 4 * THREADS * GROUPS;
  constexpr int T = CAP - 4;
  auto kern = spmv_tile_kernel<V, I, P, THREADS, GROUPS, MINB, SCALAR, DOT>;
  const size_t smem = sizeof(V) * CAP + sizeof(uint16_t) * (T + 2);
  static bool attr_done = false;  // per instantiation
  static int occ = 0;
  if (!attr_done) {
    B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2S_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) occ = 1;
    attr_done = true;
  }
  DeviceProps pr;
  if (int rc = get_props(&pr)) return rc;
  int64_t grid = a.ntiles;
  if (DOT || g_waves > 0) {
    const int waves = g_waves > 0 ? g_waves : 2;
    int64_t cap = (int64_t)pr.sm_count * occ * waves;
    if (DOT && cap > WS_MAX_PARTIALS) cap = WS_MAX_PARTIALS;
    if (grid > cap) grid = cap;
  }
  if (grid > 2147483647LL) grid = 2147483647LL;
  kern<<<(unsigned)grid, THREADS, smem, a.st>>>(a.ntiles, (const P*)a.indptr, (const I*)a.indices, (const V*)a.vals,
                                                (const V*)a.x, (V*)a.y, a.plan, a.vec_ok, (const V*)a.w,
                                                (V*)a.dot_out, a.ws);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

template <typename V, typename I, typename P, int NC, int G, int STAGES, int MINB, bool UNI, bool DOT>
static int launch_tma_u(const SpmvArgs& a) {
  using LY = TmaLayout<V, I, P, NC, G, STAGES>;
  constexpr int THREADS = (NC + 1) * 32;
  auto kern = spmv_tma_kernel<V, I, P, NC, G, STAGES, MINB, UNI, DOT>;
  const size_t smem = LY::TOTAL;
  static bool attr_done = false;
  static int occ = 0;
  if (!attr_done) {
    B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2S_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) occ = 1;
    attr_done = true;
  }
  DeviceProps pr;
  if (int rc = get_props(&pr)) return rc;
  int per_sm = occ;
  if (g_waves > 0 && g_waves < occ) per_sm = g_waves;
  TileOrder order;
  if (a.order) {
    order = *a.order;
  } else {
    order.nranges = 1; order.n_free = 1; order.n_flags = 0; order.pad = 0;
    order.lo[0] = a.tile_lo; order.hi[0] = a.tile_hi;
    for (int i = 1; i < 6; i++) { order.lo[i] = 0; order.hi[i] = 0; }
    for (int i = 0; i < 8; i++) order.flag[i] = nullptr;
    order.expect = 0; order.error = nullptr;
  }
  int64_t ntl = 0;
  for (int r = 0; r < order.nranges; r++) ntl += order.hi[r] - order.lo[r];
  int64_t grid = (int64_t)pr.sm_count * per_sm;
  if (grid > ntl) grid = ntl;
  if (DOT && grid > WS_MAX_PARTIALS) grid = WS_MAX_PARTIALS;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, THREADS, smem, a.st>>>(order, a.nrows, a.nnz, (const P*)a.indptr, (const I*)a.indices,
                                                (const V*)a.vals, (const V*)a.x, (V*)a.y, a.plan, (const V*)a.w,
                                                (V*)a.dot_out, a.ws);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

template <typename V, typename I, typename P, int NC, int G, int STAGES, int MINB, bool DOT>
static int launch_tma(const SpmvArgs& a) {
  // the register fast path is compiled in only for matrices where the plan found enough uniform tiles, so
  // the kernel of irregular matrices keeps its smaller code and register footprint
  if (a.uniform) return launch_tma_u<V, I, P, NC, G, STAGES, MINB, true, DOT>(a);
  return launch_tma_u<V, I, P, NC, G, STAGES, MINB, false, DOT>(a);
}

template <typename V, typename I, typename P, int KIND, int A, int B, int C, int D, bool DOT>
static int launch_cfg(const SpmvArgs& a) {
  if constexpr (KIND == 0) {
    return launch_ldg<V, I, P, A, B, C, (D != 0), DOT>(a);
  } else {
    // the bulk copies need 16-byte aligned array bases; otherwise fall back to the LDG kernel's scalar path
    // with the same tile size (same plan)
    return launch_tma<V, I, P, A, B, C, D, DOT>(a);
  }
}

template <typename V, typename I, typename P, bool DOT>
static int dispatch_cfg(int cfg, const SpmvArgs& a) {
#define B2S_CFG_CASE(ID, K, A, B, C, D)                                  \
  case ID:                                                               \
    if constexpr (ID == kDefaultCfgF64 || ID == kDefaultCfgF32 || ID == kScatterCfgF64 || ID == kScatterCfgF32 || \
                  (sizeof(I) == 4 && sizeof(P) == 4))                                             \
      return launch_cfg<V, I, P, K, A, B, C, D, DOT>(a);                 \
    else                                                                 \
      break;
  switch (cfg) {
    B2S_SPMV_CONFIGS(B2S_CFG_CASE)
    default: break;
  }
#undef B2S_CFG_CASE
  set_error("spmv tile config %d is not built for these index types", cfg);
  return B2S_EUNSUPPORTED;
}

template <typename V, bool DOT>
static int dispatch_idx(int it, int pt, int cfg, const SpmvArgs& a) {
  if (it == B2S_I32 && pt == B2S_I32) return dispatch_cfg<V, int32_t, int32_t, DOT>(cfg, a);
  if (it == B2S_I32 && pt == B2S_I64) return dispatch_cfg<V, int32_t, int64_t, DOT>(cfg, a);
  if (it == B2S_I64 && pt == B2S_I32) return dispatch_cfg<V, int64_t, int32_t, DOT>(cfg, a);
  if (it == B2S_I64 && pt == B2S_I64) return dispatch_cfg<V, int64_t, int64_t, DOT>(cfg, a);
  set_error("bad index type codes it=%d pt=%d", it, pt);
  return B2S_EINVAL;
}

template <typename V, typename I, typename P>
static int launch_rowgroup(int64_t nrows, int64_t nnz, const void* indptr, const void* indices, const void* vals,
                           const void* x, void* y, cudaStream_t st) {
  int gshift = 0;
  const int64_t avg = nrows > 0 ? (nnz + nrows - 1) / nrows : 0;
  while ((1 << gshift) < avg && gshift < 5) gshift++;
  if (avg <= 2) gshift = 0;
  DeviceProps pr;
  if (int rc = get_props(&pr)) return rc;
  int64_t want = ((nrows << gshift) + 255) / 256;
  int64_t cap = (int64_t)pr.sm_count * 8 * 8;
  int64_t grid = want < cap ? want : cap;
  if (grid < 1) grid = 1;
  spmv_rowgroup_kernel<V, I, P><<<(unsigned)grid, 256, 0, st>>>(nrows, (const P*)indptr, (const I*)indices,
                                                                (const V*)vals, (const V*)x, (V*)y, gshift);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

template <typename V>
static int dispatch_rowgroup(int it, int pt, int64_t nrows, int64_t nnz, const void* indptr, const void* indices,
                             const void* vals, const void* x, void* y, cudaStream_t st) {
  if (it == B2S_I32 && pt == B2S_I32) return launch_rowgroup<V, int32_t, int32_t>(nrows, nnz, indptr, indices, vals, x, y, st);
  if (it == B2S_I32 && pt == B2S_I64) return launch_rowgroup<V, int32_t, int64_t>(nrows, nnz, indptr, indices, vals, x, y, st);
  if (it == B2S_I64 && pt == B2S_I32) return launch_rowgroup<V, int64_t, int32_t>(nrows, nnz, indptr, indices, vals, x, y, st);
  if (it == B2S_I64 && pt == B2S_I64) return launch_rowgroup<V, int64_t, int64_t>(nrows, nnz, indptr, indices, vals, x, y, st);
  set_error("bad index type codes it=%d pt=%d", it, pt);
  return B2S_EINVAL;
}

static int check_common(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                        const void* indices, const void* vals, const void* x, const void* y) {
  B2S_CHECK_ARG(vt == B2S_F32 || vt == B2S_F64, "bad value type code %d", vt);
  B2S_CHECK_ARG(it == B2S_I32 || it == B2S_I64, "bad index type code %d", it);
  B2S_CHECK_ARG(pt == B2S_I32 || pt == B2S_I64, "bad indptr type code %d", pt);
  B2S_CHECK_ARG(nrows >= 0 && ncols >= 0 && nnz >= 0, "negative dimension");
  B2S_CHECK_ARG(nrows < 2147483647LL, "nrows >= 2^31-1 is not supported");
  B2S_CHECK_ARG(indptr != nullptr, "indptr is NULL");
  B2S_CHECK_ARG(nnz == 0 || (indices != nullptr && vals != nullptr), "indices/vals NULL with nnz > 0");
  B2S_CHECK_ARG(ncols == 0 || nnz == 0 || x != nullptr, "x is NULL");
  B2S_CHECK_ARG(nrows == 0 || y != nullptr, "y is NULL");
  B2S_CHECK_ARG(pt == B2S_I64 || nnz < 2147483647LL, "int32 indptr cannot address nnz >= 2^31-1");
  return B2S_OK;
}

}  // namespace b2s

namespace b2s {

// Column-locality statistic: mean number of distinct 128-byte lines of x touched by 32 consecutive
// nonzeros (one warp-wide gather), sampled at up to 4096 evenly spaced positions.  ~5 for stencils,
// 32 for uniformly random columns.  Scattered matrices are L1-tag-bound and want many resident warps,
// so they are routed to the row-group kernel instead of the staged-tile kernel.
template <typename I>
__global__ void __launch_bounds__(256)
spmv_locality_kernel(int64_t nnz, const I* __restrict__ indices, int elem_shift, int64_t nsamples,
                     unsigned long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= nsamples) return;
  const int64_t p = (nnz - 32) * w / nsamples;
  const long long line = ((long long)indices[p + lane] << elem_shift) >> 7;
  const unsigned m = __match_any_sync(0xffffffffu, line);
  const unsigned leaders = __ballot_sync(0xffffffffu, (__ffs(m) - 1) == lane);
  if (lane == 0) atomicAdd(out, (unsigned long long)__popc(leaders));
}

// [min col, max col] of the nonzeros [k_lo, k_hi) -- the x window a chunk of rows reads (reference:
// MinMaxImagePartition, sparse/partition.py:139-208 / bounds_from_partitioned_coordinates.cu:35-62).
template <typename I>
__global__ void __launch_bounds__(256)
spmv_col_window_kernel(const I* __restrict__ indices, const long long* __restrict__ kbounds, int nchunks,
                       int blocks_per_chunk, long long* __restrict__ out_minmax) {
  const int c = blockIdx.x / blocks_per_chunk, b = blockIdx.x % blocks_per_chunk;
  if (c >= nchunks) return;
  const long long lo = kbounds[c], hi = kbounds[c + 1];
  long long mn = LLONG_MAX, mx = -1;
  for (long long k = lo + (long long)b * 256 + threadIdx.x; k < hi; k += 256LL * blocks_per_chunk) {
    const long long v = (long long)indices[k];
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, mn, o), bb = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = bb > mx ? bb : mx;
  }
  if ((threadIdx.x & 31) == 0 && mx >= 0) {
    atomicMin(reinterpret_cast<long long*>(&out_minmax[2 * c]), mn);
    atomicMax(reinterpret_cast<long long*>(&out_minmax[2 * c + 1]), mx);
  }
}

constexpr int kPlanChunks = 16;

struct PlanHandle {
  uint32_t magic;
  int vt, it, pt;
  int cfg;          // tile configuration the device plan was built for
  int use_rowgroup; // 1: matrix judged scattered -> plan-free row-group kernel
  int use_uniform;  // 1: >= 25% of the tiles are ELL-like -> kernel variant with the register fast path
  int scattered;    // 1: > 16 distinct x lines per warp-wide gather
  int64_t nrows, ncols, nnz, ntiles;
  double lines_per_warp;
  const PlanEntry* dev;
  // row chunks for pipelined host<->device SpMV: chunk c = tiles [ctile[c], ctile[c+1]), rows [crow[c], crow[c+1]),
  // reading x only inside [ccol_lo[c], ccol_hi[c])
  int nchunks;
  int64_t ctile[kPlanChunks + 1], crow[kPlanChunks + 1], ccol_lo[kPlanChunks], ccol_hi[kPlanChunks];
};
static constexpr uint32_t kPlanMagic = 0xB2005A17u;

}  // namespace b2s

using namespace b2s;

extern "C" {

// Debug/tuning hooks (not part of the documented ABI; used by tools/ sweep scripts).
// cfg < 0 restores automatic selection.
int b2s_spmv_set_config(int cfg, int waves) {
  if (cfg >= kNumCfgs) { set_error("config %d out of range [0,%d)", cfg, kNumCfgs); return B2S_EINVAL; }
  g_cfg = cfg < 0 ? -1 : cfg;
  g_waves = waves < 0 ? 0 : waves;
  return B2S_OK;
}
int b2s_spmv_get_config(void) { return g_cfg; }
int b2s_spmv_num_configs(void) { return kNumCfgs; }

static int64_t tiles_for(int cfg, int vt, int64_t nrows, int64_t nnz) {
  if (nrows <= 0 || nnz < 0) return 0;
  const int64_t T = cfg_T(cfg, vt);
  return (nrows + nnz + T - 1) / T;
}

int64_t b2s_spmv_plan_tiles(int vt, int64_t nrows, int64_t nnz) {
  // upper bound over the configurations plan_create may pick (it decides after sampling the matrix)
  const int64_t a = tiles_for(resolve_cfg(vt, false), vt, nrows, nnz);
  const int64_t b = tiles_for(resolve_cfg(vt, true), vt, nrows, nnz);
  return a > b ? a : b;
}

int64_t b2s_spmv_plan_bytes(int vt, int64_t nrows, int64_t nnz) {
  // (ntiles + 1) entries + one 16-byte slot for the plan statistics
  return (b2s_spmv_plan_tiles(vt, nrows, nnz) + 2) * (int64_t)sizeof(PlanEntry);
}

int b2s_spmv_plan_create(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                         const void* indices, void* plan_buf, void* stream, void** plan_out) {
  B2S_CHECK_ARG(plan_out != nullptr, "plan_out is NULL");
  *plan_out = nullptr;
  if (int rc = check_common(vt, it, pt, nrows, ncols, nnz, indptr, indices, indices, plan_out, plan_out)) return rc;
  B2S_CHECK_ARG(plan_buf != nullptr && aligned16(plan_buf), "plan buffer must be non-NULL and 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  PlanEntry* dev = (PlanEntry*)plan_buf;
  // the statistics slot sits after the largest plan this buffer may hold
  unsigned long long* stat = (unsigned long long*)(dev + b2s_spmv_plan_tiles(vt, nrows, nnz) + 1);
  double lines = 0.0;
  int64_t uniform_tiles = 0;
  // 1. column locality (decides the kernel family / tile shape)
  if (nrows > 0 && nnz >= 64) {
    int64_t ns = nnz / 32;
    if (ns > 4096) ns = 4096;
    B2S_CUDA(cudaMemsetAsync(stat, 0, 16, st));
    const unsigned g2 = (unsigned)((ns * 32 + 255) / 256);
    const int shift = vt == B2S_F32 ? 2 : 3;
    if (it == B2S_I32) spmv_locality_kernel<int32_t><<<g2, 256, 0, st>>>(nnz, (const int32_t*)indices, shift, ns, stat);
    else               spmv_locality_kernel<int64_t><<<g2, 256, 0, st>>>(nnz, (const int64_t*)indices, shift, ns, stat);
    B2S_LAUNCH_CHECK();
    unsigned long long total = 0;
    B2S_CUDA(cudaMemcpyAsync(&total, stat, sizeof(total), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    lines = (double)total / (double)ns;
  }
  const bool scattered = lines > 16.0;
  const int cfg = resolve_cfg(vt, scattered);
  const int64_t ntiles = tiles_for(cfg, vt, nrows, nnz);
  // 2. tile boundaries + uniform-row annotation for the chosen tile shape
  if (ntiles > 0) {
    const int64_t T = cfg_T(cfg, vt);
    const unsigned grid = (unsigned)((ntiles + 1 + 255) / 256);
    if (pt == B2S_I32) spmv_plan_kernel<int32_t><<<grid, 256, 0, st>>>(nrows, (const int32_t*)indptr, T, ntiles, dev);
    else               spmv_plan_kernel<int64_t><<<grid, 256, 0, st>>>(nrows, (const int64_t*)indptr, T, ntiles, dev);
    B2S_LAUNCH_CHECK();
    B2S_CUDA(cudaMemsetAsync(stat, 0, 16, st));
    const unsigned gu = (unsigned)((ntiles + 255) / 256);
    const int ept = vt == B2S_F32 ? 4 : 2;
    if (pt == B2S_I32) spmv_plan_uniform_kernel<int32_t><<<gu, 256, 0, st>>>((const int32_t*)indptr, ntiles, dev, ept, stat + 1);
    else               spmv_plan_uniform_kernel<int64_t><<<gu, 256, 0, st>>>((const int64_t*)indptr, ntiles, dev, ept, stat + 1);
    B2S_LAUNCH_CHECK();
    unsigned long long ut = 0;
    B2S_CUDA(cudaMemcpyAsync(&ut, stat + 1, sizeof(ut), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    uniform_tiles = (int64_t)ut;
  }
  PlanHandle* h = new PlanHandle();
  h->magic = kPlanMagic;
  h->vt = vt; h->it = it; h->pt = pt; h->cfg = cfg;
  h->use_rowgroup = 0;  // scattered matrices now get a deep-MLP tile shape instead (see kScatterCfg*)
  h->scattered = scattered ? 1 : 0;
  // kernel flavour: the variant with the register fast path and the conflict-avoiding skewed reduce is used
  // when >= 25% of the tiles are ELL-like, or rows are long (>= 12) or of even mean length; short odd rows
  // (e.g. 5-point stencils) keep the lean sequential-reduce variant
  const int64_t meanL = nrows > 0 ? (nnz + nrows / 2) / nrows : 0;
  h->use_uniform = (ntiles > 0 && (uniform_tiles * 4 >= ntiles || meanL >= 12 || (meanL > 0 && (meanL & 1) == 0))) ? 1 : 0;
  h->nrows = nrows; h->ncols = ncols; h->nnz = nnz; h->ntiles = ntiles;
  h->lines_per_warp = lines;
  h->dev = dev;
  h->nchunks = 0;
  if (ntiles >= 4 * kPlanChunks && nnz > 0) {
    // chunk boundaries (tiles), their rows / nnz ranges (read back from the device plan) and column windows
    const int K = kPlanChunks;
    PlanEntry ends[kPlanChunks + 1];
    for (int c = 0; c <= K; c++) {
      h->ctile[c] = ntiles * c / K;
      B2S_CUDA(cudaMemcpyAsync(&ends[c], dev + h->ctile[c], sizeof(PlanEntry), cudaMemcpyDeviceToHost, st));
    }
    B2S_CUDA(cudaStreamSynchronize(st));
    long long kb_host[kPlanChunks + 1], mm_host[2 * kPlanChunks];
    for (int c = 0; c <= K; c++) { h->crow[c] = ends[c].row; kb_host[c] = ends[c].k; }
    for (int c = 0; c < K; c++) { mm_host[2 * c] = LLONG_MAX; mm_host[2 * c + 1] = -1; }
    // scratch: the tail of the plan buffer is free to borrow here?  No -- use a small temporary allocation.
    long long* dtmp = nullptr;
    B2S_CUDA(cudaMallocAsync((void**)&dtmp, sizeof(long long) * (3 * K + 1), st));
    B2S_CUDA(cudaMemcpyAsync(dtmp, kb_host, sizeof(long long) * (K + 1), cudaMemcpyHostToDevice, st));
    B2S_CUDA(cudaMemcpyAsync(dtmp + K + 1, mm_host, sizeof(long long) * 2 * K, cudaMemcpyHostToDevice, st));
    const int bpc = 64;
    if (it == B2S_I32) spmv_col_window_kernel<int32_t><<<K * bpc, 256, 0, st>>>((const int32_t*)indices, dtmp, K, bpc, dtmp + K + 1);
    else               spmv_col_window_kernel<int64_t><<<K * bpc, 256, 0, st>>>((const int64_t*)indices, dtmp, K, bpc, dtmp + K + 1);
    B2S_LAUNCH_CHECK();
    B2S_CUDA(cudaMemcpyAsync(mm_host, dtmp + K + 1, sizeof(long long) * 2 * K, cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    B2S_CUDA(cudaFreeAsync(dtmp, st));
    for (int c = 0; c < K; c++) {
      if (mm_host[2 * c + 1] < 0) { h->ccol_lo[c] = 0; h->ccol_hi[c] = 0; }
      else { h->ccol_lo[c] = mm_host[2 * c]; h->ccol_hi[c] = mm_host[2 * c + 1] + 1; }
    }
    h->nchunks = K;
  }
  *plan_out = h;
  return B2S_OK;
}

/* Row chunks of a plan for pipelined host<->device products: out = nchunks x {tile_lo, tile_hi, row_lo, row_hi,
 * col_lo, col_hi}; returns the number of chunks through *nchunks_host (0: plan too small / not chunked). */
int b2s_spmv_plan_chunks(const void* plan, int64_t* out_host, int max_chunks, int* nchunks_host) {
  const PlanHandle* h = (const PlanHandle*)plan;
  B2S_CHECK_ARG(h && h->magic == kPlanMagic && out_host && nchunks_host, "bad plan handle / out pointer");
  const int n = h->nchunks < max_chunks ? h->nchunks : max_chunks;
  for (int c = 0; c < n; c++) {
    out_host[6 * c + 0] = h->ctile[c]; out_host[6 * c + 1] = h->ctile[c + 1];
    out_host[6 * c + 2] = h->crow[c];  out_host[6 * c + 3] = h->crow[c + 1];
    out_host[6 * c + 4] = h->ccol_lo[c]; out_host[6 * c + 5] = h->ccol_hi[c];
  }
  *nchunks_host = (kCfgs[h->cfg].kind == 1 && !h->use_rowgroup) ? n : 0;
  return B2S_OK;
}

int b2s_spmv_plan_destroy(void* plan) {
  PlanHandle* h = (PlanHandle*)plan;
  if (!h) return B2S_OK;
  B2S_CHECK_ARG(h->magic == kPlanMagic, "not a b2s spmv plan handle");
  h->magic = 0;
  delete h;
  return B2S_OK;
}

/* out[0] = tile config, out[1] = 1 if the row-group kernel is selected, out[2] = ntiles,
 * out[3] = 1000 * mean distinct x lines per 32 consecutive nonzeros */
int b2s_spmv_plan_info(const void* plan, int64_t* out4_host) {
  const PlanHandle* h = (const PlanHandle*)plan;
  B2S_CHECK_ARG(h && h->magic == kPlanMagic && out4_host, "bad plan handle / out pointer");
  out4_host[0] = h->cfg; out4_host[1] = h->use_rowgroup + 2 * h->use_uniform + 4 * h->scattered; out4_host[2] = h->ntiles;
  out4_host[3] = (int64_t)(h->lines_per_warp * 1000.0);
  return B2S_OK;
}

/* force the kernel family for a plan: 0 = staged tiles, 1 = row-group (tools / tests) */
int b2s_spmv_plan_set_kernel(void* plan, int use_rowgroup) {
  PlanHandle* h = (PlanHandle*)plan;
  B2S_CHECK_ARG(h && h->magic == kPlanMagic, "bad plan handle");
  h->use_rowgroup = use_rowgroup ? 1 : 0;
  return B2S_OK;
}

static int spmv_impl(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y, const void* w, void* dot_out,
                     const void* plan, void* ws, void* stream, bool dot, int64_t tile_lo = 0, int64_t tile_hi = -1,
                     const TileOrder* order = nullptr) {
  if (int rc = check_common(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dot) {
    B2S_CHECK_ARG(ws != nullptr && dot_out != nullptr, "ws/dot_out is NULL");
    B2S_CHECK_ARG(nrows == 0 || w != nullptr, "w is NULL");
  }
  if (nrows == 0) {
    if (dot) B2S_CUDA(cudaMemsetAsync(dot_out, 0, vt == B2S_F32 ? 4 : 8, st));
    return B2S_OK;
  }
  const PlanHandle* h = (const PlanHandle*)plan;
  if (h) {
    B2S_CHECK_ARG(h->magic == kPlanMagic, "plan is not a handle from b2s_spmv_plan_create");
    B2S_CHECK_ARG(h->vt == vt && h->it == it && h->pt == pt && h->nrows == nrows && h->ncols == ncols && h->nnz == nnz,
                  "plan was created for a different matrix (types or dimensions differ)");
  }
  const bool aligned = (nnz == 0) || (aligned16(indices) && aligned16(vals));
  bool rowgroup = (h == nullptr) || h->use_rowgroup;
  if (h && !rowgroup && kCfgs[h->cfg].kind == 1 && !(aligned && aligned16(indptr))) rowgroup = true;  // TMA needs 16-byte aligned bases
  B2S_CHECK_ARG(!(rowgroup && order), "the fused halo SpMV needs the TMA tile kernel (aligned arrays, tile plan)");
  if (rowgroup) {
    int rc = (vt == B2S_F32) ? dispatch_rowgroup<float>(it, pt, nrows, nnz, indptr, indices, vals, x, y, st)
                             : dispatch_rowgroup<double>(it, pt, nrows, nnz, indptr, indices, vals, x, y, st);
    if (rc || !dot) return rc;
    return b2s_dot(vt, nrows, w, y, dot_out, ws, stream);
  }
  SpmvArgs a;
  a.ntiles = h->ntiles;
  a.tile_lo = tile_lo;
  a.tile_hi = tile_hi < 0 ? h->ntiles : tile_hi;
  a.order = order;
  B2S_CHECK_ARG(order == nullptr || kCfgs[h->cfg].kind == 1, "explicit tile orders need a TMA tile plan");
  B2S_CHECK_ARG(a.tile_lo >= 0 && a.tile_lo <= a.tile_hi && a.tile_hi <= h->ntiles, "tile range out of bounds");
  B2S_CHECK_ARG((a.tile_lo == 0 && a.tile_hi == h->ntiles) || kCfgs[h->cfg].kind == 1,
                "tile sub-ranges need a TMA tile plan");
  if (a.tile_lo == a.tile_hi) return B2S_OK;
  a.nrows = nrows; a.nnz = nnz;
  a.indptr = indptr; a.indices = indices; a.vals = vals; a.x = x; a.y = y;
  a.plan = h->dev;
  a.vec_ok = aligned ? 1 : 0;
  a.uniform = h->use_uniform;
  a.w = dot ? w : nullptr; a.dot_out = dot ? dot_out : nullptr; a.ws = dot ? ws : nullptr;
  a.st = st;
  if (vt == B2S_F32) {
    if (dot) return dispatch_idx<float, true>(it, pt, h->cfg, a);
    return dispatch_idx<float, false>(it, pt, h->cfg, a);
  }
  if (dot) return dispatch_idx<double, true>(it, pt, h->cfg, a);
  return dispatch_idx<double, false>(it, pt, h->cfg, a);
}

int b2s_spmv_csr(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                 const void* indices, const void* vals, const void* x, void* y, const void* plan, void* stream) {
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, nullptr, nullptr, plan, nullptr,
                   stream, false);
}

/* y[rows of tiles [tile_lo, tile_hi)] = (A x) restricted to those rows; needs x valid on the chunk's column window */
int b2s_spmv_csr_tiles(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                       const void* indices, const void* vals, const void* x, void* y, const void* plan,
                       int64_t tile_lo, int64_t tile_hi, void* stream) {
  B2S_CHECK_ARG(plan != nullptr, "b2s_spmv_csr_tiles needs a plan");
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, nullptr, nullptr, plan, nullptr,
                   stream, false, tile_lo, tile_hi);
}

/* SpMV fused with the arrival of a halo that peer GPUs push into x (b2s_peer_halo_push): tiles are visited range by
 * range -- `ranges_host` = nranges x {tile_lo, tile_hi}, the first n_free ranges read only locally valid x --
 * and before its first non-free tile each CTA polls the nflags arrival flags (device addresses in this GPU's
 * memory) until they reach `expect`.  On timeout *error_flag_dev is set to 1 and the product continues. */
int b2s_spmv_csr_halo(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                      const void* indices, const void* vals, const void* x, void* y, const void* plan, int nranges,
                      const int64_t* ranges_host, int n_free, int nflags, void* const* flag_ptrs_host,
                      uint64_t expect, void* error_flag_dev, void* stream) {
  B2S_CHECK_ARG(plan != nullptr, "b2s_spmv_csr_halo needs a plan");
  B2S_CHECK_ARG(nranges >= 1 && nranges <= 6 && ranges_host && n_free >= 0 && n_free <= nranges, "bad tile ranges");
  B2S_CHECK_ARG(nflags >= 0 && nflags <= 8 && (nflags == 0 || (flag_ptrs_host && error_flag_dev)), "bad flag list");
  TileOrder o;
  o.nranges = nranges; o.n_free = n_free; o.n_flags = nflags; o.pad = 0;
  for (int i = 0; i < 6; i++) { o.lo[i] = i < nranges ? ranges_host[2 * i] : 0; o.hi[i] = i < nranges ? ranges_host[2 * i + 1] : 0; }
  for (int i = 0; i < nranges; i++) B2S_CHECK_ARG(o.lo[i] >= 0 && o.hi[i] >= o.lo[i], "bad tile range %d", i);
  for (int i = 0; i < 8; i++) o.flag[i] = i < nflags ? (const unsigned long long*)flag_ptrs_host[i] : nullptr;
  o.expect = expect;
  o.error = (unsigned long long*)error_flag_dev;
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, nullptr, nullptr, plan, nullptr,
                   stream, false, 0, -1, &o);
}

/* y_host = A x_host with HOST vectors (matrix resident on the device): x is streamed in and y streamed out chunk
 * by chunk on two internal copy streams while the tiles of each chunk run on `stream`, so both PCIe directions
 * and the kernel overlap (see b2s_spmv_plan_chunks).  x_dev / y_dev are caller-owned device scratch vectors of
 * ncols / nrows elements.  Pinned host memory gives true overlap.  Returns after y_host is complete (syncs). */
int b2s_spmv_csr_host(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                      const void* indices, const void* vals, const void* x_host, void* y_host, void* x_dev,
                      void* y_dev, const void* plan, void* stream) {
  B2S_CHECK_ARG(vt == B2S_F32 || vt == B2S_F64, "bad value type code %d", vt);
  B2S_CHECK_ARG((ncols == 0 || (x_host && x_dev)) && (nrows == 0 || (y_host && y_dev)), "NULL vector pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sv = vt == B2S_F32 ? 4 : 8;
  const PlanHandle* h = (const PlanHandle*)plan;
  const bool chunked = h && h->magic == kPlanMagic && h->nchunks > 0 && kCfgs[h->cfg].kind == 1 && !h->use_rowgroup;
  if (!chunked) {
    B2S_CUDA(cudaMemcpyAsync(x_dev, x_host, sv * (size_t)ncols, cudaMemcpyHostToDevice, st));
    if (int rc = b2s_spmv_csr(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x_dev, y_dev, plan, stream)) return rc;
    B2S_CUDA(cudaMemcpyAsync(y_host, y_dev, sv * (size_t)nrows, cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    return B2S_OK;
  }
  // per-device copy streams and events, created once
  struct Pipe { cudaStream_t s_in = nullptr, s_out = nullptr; cudaEvent_t ev_in[kPlanChunks], ev_k[kPlanChunks], ev0; bool ok = false; };
  static Pipe pipes[64];
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  B2S_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal out of range");
  Pipe& P = pipes[dev];
  if (!P.ok) {
    B2S_CUDA(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
    B2S_CUDA(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
    for (int c = 0; c < kPlanChunks; c++) {
      B2S_CUDA(cudaEventCreateWithFlags(&P.ev_in[c], cudaEventDisableTiming));
      B2S_CUDA(cudaEventCreateWithFlags(&P.ev_k[c], cudaEventDisableTiming));
    }
    B2S_CUDA(cudaEventCreateWithFlags(&P.ev0, cudaEventDisableTiming));
    P.ok = true;
  }
  // the copy streams start after everything already queued on the compute stream
  B2S_CUDA(cudaEventRecord(P.ev0, st));
  B2S_CUDA(cudaStreamWaitEvent(P.s_in, P.ev0, 0));
  B2S_CUDA(cudaStreamWaitEvent(P.s_out, P.ev0, 0));
  int64_t copied = 0;
  for (int c = 0; c < h->nchunks; c++) {
    const int64_t need = h->ccol_hi[c];
    if (need > copied) {
      B2S_CUDA(cudaMemcpyAsync((char*)x_dev + sv * copied, (const char*)x_host + sv * copied, sv * (size_t)(need - copied),
                               cudaMemcpyHostToDevice, P.s_in));
      copied = need;
      B2S_CUDA(cudaEventRecord(P.ev_in[c], P.s_in));
      B2S_CUDA(cudaStreamWaitEvent(st, P.ev_in[c], 0));
    }
    if (int rc = b2s_spmv_csr_tiles(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x_dev, y_dev, plan,
                                    h->ctile[c], h->ctile[c + 1], stream)) return rc;
    const int64_t r0 = h->crow[c], r1 = h->crow[c + 1];
    if (r1 > r0) {
      B2S_CUDA(cudaEventRecord(P.ev_k[c], st));
      B2S_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_k[c], 0));
      B2S_CUDA(cudaMemcpyAsync((char*)y_host + sv * r0, (const char*)y_dev + sv * r0, sv * (size_t)(r1 - r0),
                               cudaMemcpyDeviceToHost, P.s_out));
    }
  }
  B2S_CUDA(cudaStreamSynchronize(P.s_out));
  B2S_CUDA(cudaStreamSynchronize(st));
  return B2S_OK;
}

int b2s_spmv_csr_dot(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y, const void* w, void* dot_out,
                     const void* plan, void* ws, void* stream) {
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, w, dot_out, plan, ws, stream, true);
}

}  // extern "C"
