// spmv.cu -- CSR SpMV for sm_100a.
//
// Replaces CSRSpMVRowSplit::gpu_variant (reference src/sparse/array/csr/spmv.cu:24-123, a
// cusparseSpMV call) with a hand-written row-blocked ("merge-path tiles rounded to row
// boundaries") kernel:
//
//   plan   : the (rows + nnz) work list is cut into tiles of T merge items; tile t owns the rows
//            whose start position indptr[r] + r falls in [t*T, (t+1)*T).  Every tile therefore has
//            <= T rows, and all of its rows except possibly the last fit in one shared-memory
//            chunk of CAP = T + 4 nonzeros.  The plan is (ntiles + 1) int32 row boundaries.
//   phase A: the CTA streams the tile's contiguous nnz range with 128-bit evict-first loads of
//            indices/vals, gathers x through the read-only path and parks vals[k]*x[col[k]] in
//            shared memory (fully coalesced regardless of row lengths).
//   tail   : a last row longer than the chunk is finished by the whole CTA straight from global
//            memory (block reduction) -- so no cross-CTA carries, no atomics, no fix-up pass.
//   phase B: 2^s lanes per row (s chosen per tile from its mean row length) reduce the parked
//            products: sequential for short rows (same order as the reference's CPU loop,
//            spmv.cc:36-44), warp-shuffle tree for long rows.  Optional fused epilogue: the CG
//            inner product sum_i w[i]*y[i] (deterministic two-stage grid reduction).
//
// HBM-bound by construction: algorithmic bytes per launch are
//   nnz*(sizeof V + sizeof I) + (nrows+1)*sizeof P + ncols*sizeof V + nrows*sizeof V.
#include "common.cuh"

namespace b2s {

// ---------------------------------------------------------------------------------------------
// Tile configurations.  CAP = 4*THREADS*GROUPS nonzeros staged per chunk; T = CAP - 4 merge items.
// ---------------------------------------------------------------------------------------------
// MINB = CTAs/SM promised to the compiler (register cap = 65536 / (THREADS*MINB)).
// X(ID, THREADS, GROUPS, MINB).  Config 0 is the default and the only one instantiated for
// int64 index/indptr types; the others exist for tuning sweeps (tools/, b2s_spmv_set_config).
#define B2S_SPMV_CONFIGS(X) \
  X(0, 256, 4, 3)           \
  X(1, 256, 2, 4)           \
  X(2, 128, 4, 6)           \
  X(3, 256, 4, 4)           \
  X(4, 256, 4, 2)           \
  X(5, 512, 2, 2)           \
  X(6, 256, 8, 2)           \
  X(7, 128, 8, 4)
struct TileCfgRt { int threads, groups, minb; };
static const TileCfgRt kCfgs[] = {
#define X(ID, TH, GR, MB) {TH, GR, MB},
    B2S_SPMV_CONFIGS(X)
#undef X
};
static constexpr int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);
static int g_cfg = 0;          // selected configuration (b2s_spmv_set_config)
static int g_waves = 0;        // 0 = one tile per CTA; >0 = grid-stride with waves*SMs*occupancy CTAs

static inline int cfg_cap(int c) { return 4 * kCfgs[c].threads * kCfgs[c].groups; }
static inline int cfg_T(int c) { return cfg_cap(c) - 4; }

// ---------------------------------------------------------------------------------------------
// Plan kernel: plan[t] = first row r with indptr[r] + r >= t*T ; plan[ntiles] = nrows.
// ---------------------------------------------------------------------------------------------
template <typename P>
__global__ void spmv_plan_kernel(int64_t nrows, const P* __restrict__ indptr, int64_t T, int64_t ntiles,
                                 int32_t* __restrict__ plan) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  if (t == ntiles) { plan[t] = (int32_t)nrows; return; }
  const int64_t target = t * T;
  int64_t lo = 0, hi = nrows;  // answer in [0, nrows]
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)indptr[mid] + mid >= target) hi = mid; else lo = mid + 1;
  }
  plan[t] = (int32_t)lo;
}

// ---------------------------------------------------------------------------------------------
// Tiled SpMV kernel.
// ---------------------------------------------------------------------------------------------
template <typename V, typename I, typename P, int THREADS, int GROUPS, int MINB, bool DOT>
__global__ void __launch_bounds__(THREADS, MINB)
spmv_tile_kernel(int64_t ntiles, const P* __restrict__ indptr, const I* __restrict__ indices,
                 const V* __restrict__ vals, const V* __restrict__ x, V* __restrict__ y,
                 const int32_t* __restrict__ plan, int vec_ok, const V* __restrict__ w, V* dot_out, void* ws) {
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
    const int r0 = plan[t], r1 = plan[t + 1];
    const int nr = r1 - r0;
    if (nr <= 0) continue;  // block-uniform: tile lies inside a long row owned by an earlier tile
    const int64_t k0 = (int64_t)indptr[r0];
    const int64_t k1 = (int64_t)indptr[r1];
    const int64_t kb = k0 & ~(int64_t)3;           // 16-byte aligned chunk base
    const int off = (int)(k0 - kb);
    const int64_t kce = (k1 < kb + CAP) ? k1 : kb + CAP;  // end of the staged chunk
    const bool has_tail = k1 > kce;

    __syncthreads();  // previous tile's phase B is done with prod/sptr

    // row offsets of this tile, relative to k0, clamped to the chunk (uint16: CAP <= 32768)
    for (int j = tid; j <= nr; j += THREADS) {
      int64_t rel = (int64_t)indptr[r0 + j] - k0;
      int64_t lim = kce - k0;
      sptr[j] = (uint16_t)(rel < lim ? rel : lim);
    }

    // ---- phase A: stream nnz [k0, kce) -> prod[k - kb] ---------------------------------------
    if (vec_ok) {
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
#pragma unroll
          for (int q = 0; q < 4; q++) prod[e + q] = a[g][q] * xv[g][q];
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

    // ---- phase B: per-row reduction of the parked products ----------------------------------------
    // lanes per row: 1 for short rows (sequential, reference order), else pow2 >= mean row length
    const int64_t nnz_t = k1 - k0;
    int gshift = 0;
    if (nnz_t > 6 * (int64_t)nr) {
      const int avg = (int)((nnz_t + nr - 1) / nr);
      while ((1 << gshift) < avg && gshift < 5) gshift++;
    }
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
      V sum = 0;
      for (int k = s + lig; k < e; k += g) sum += pr[k];
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
template <typename V, typename I, typename P, int THREADS, int GROUPS, int MINB, bool DOT>
static int launch_tile(int64_t ntiles, const void* indptr, const void* indices, const void* vals,
                       const void* x, void* y, const int32_t* plan, int vec_ok, const void* w, void* dot_out,
                       void* ws, cudaStream_t st) {
  constexpr int CAP = 4 * THREADS * GROUPS;
  constexpr int T = CAP - 4;
  auto kern = spmv_tile_kernel<V, I, P, THREADS, GROUPS, MINB, DOT>;
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
  int64_t grid = ntiles;
  if (DOT || g_waves > 0) {
    const int waves = g_waves > 0 ? g_waves : 4;
    int64_t cap = (int64_t)pr.sm_count * occ * waves;
    if (DOT && cap > WS_MAX_PARTIALS) cap = WS_MAX_PARTIALS;
    if (grid > cap) grid = cap;
  }
  if (grid > 2147483647LL) grid = 2147483647LL;
  kern<<<(unsigned)grid, THREADS, smem, st>>>(ntiles, (const P*)indptr, (const I*)indices, (const V*)vals,
                                              (const V*)x, (V*)y, plan, vec_ok, (const V*)w, (V*)dot_out, ws);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

template <typename V, typename I, typename P, bool DOT>
static int dispatch_cfg(int cfg, int64_t ntiles, const void* indptr, const void* indices, const void* vals,
                        const void* x, void* y, const int32_t* plan, int vec_ok, const void* w, void* dot_out,
                        void* ws, cudaStream_t st) {
#define B2S_CFG_CASE(ID, TH, GR, MB)                                                                     \
  case ID:                                                                                               \
    if constexpr (ID == 0 || (sizeof(I) == 4 && sizeof(P) == 4))                                         \
      return launch_tile<V, I, P, TH, GR, MB, DOT>(ntiles, indptr, indices, vals, x, y, plan, vec_ok, w, \
                                                   dot_out, ws, st);                                     \
    else                                                                                                 \
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
static int dispatch_idx(int it, int pt, int cfg, int64_t ntiles, const void* indptr, const void* indices,
                        const void* vals, const void* x, void* y, const int32_t* plan, int vec_ok,
                        const void* w, void* dot_out, void* ws, cudaStream_t st) {
  if (it == B2S_I32 && pt == B2S_I32) return dispatch_cfg<V, int32_t, int32_t, DOT>(cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, w, dot_out, ws, st);
  if (it == B2S_I32 && pt == B2S_I64) return dispatch_cfg<V, int32_t, int64_t, DOT>(cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, w, dot_out, ws, st);
  if (it == B2S_I64 && pt == B2S_I32) return dispatch_cfg<V, int64_t, int32_t, DOT>(cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, w, dot_out, ws, st);
  if (it == B2S_I64 && pt == B2S_I64) return dispatch_cfg<V, int64_t, int64_t, DOT>(cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, w, dot_out, ws, st);
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

using namespace b2s;

extern "C" {

// Debug/tuning hooks (not part of the documented ABI; used by tools/sweep scripts).
int b2s_spmv_set_config(int cfg, int waves) {
  if (cfg < 0 || cfg >= kNumCfgs) { set_error("config %d out of range [0,%d)", cfg, kNumCfgs); return B2S_EINVAL; }
  g_cfg = cfg;
  g_waves = waves < 0 ? 0 : waves;
  return B2S_OK;
}
int b2s_spmv_get_config(void) { return g_cfg; }

int64_t b2s_spmv_plan_tiles(int vt, int64_t nrows, int64_t nnz) {
  (void)vt;
  if (nrows <= 0 || nnz < 0) return 0;
  const int64_t T = cfg_T(g_cfg);
  return (nrows + nnz + T - 1) / T;
}

int b2s_spmv_plan_build(int vt, int pt, int64_t nrows, int64_t nnz, const void* indptr, int32_t* plan,
                        void* stream) {
  B2S_CHECK_ARG(pt == B2S_I32 || pt == B2S_I64, "bad indptr type code %d", pt);
  B2S_CHECK_ARG(nrows >= 0 && nnz >= 0, "negative dimension");
  B2S_CHECK_ARG(nrows < 2147483647LL, "nrows >= 2^31-1 is not supported");
  const int64_t ntiles = b2s_spmv_plan_tiles(vt, nrows, nnz);
  if (ntiles == 0) return B2S_OK;
  B2S_CHECK_ARG(indptr != nullptr && plan != nullptr, "indptr/plan is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t T = cfg_T(g_cfg);
  const unsigned grid = (unsigned)((ntiles + 1 + 255) / 256);
  if (pt == B2S_I32) spmv_plan_kernel<int32_t><<<grid, 256, 0, st>>>(nrows, (const int32_t*)indptr, T, ntiles, plan);
  else               spmv_plan_kernel<int64_t><<<grid, 256, 0, st>>>(nrows, (const int64_t*)indptr, T, ntiles, plan);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

static int spmv_impl(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y, const void* w, void* dot_out,
                     const int32_t* plan, void* ws, void* stream, bool dot) {
  if (int rc = check_common(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dot) {
    B2S_CHECK_ARG(plan != nullptr, "b2s_spmv_csr_dot requires a plan");
    B2S_CHECK_ARG(ws != nullptr && dot_out != nullptr, "ws/dot_out is NULL");
    B2S_CHECK_ARG(nrows == 0 || w != nullptr, "w is NULL");
  }
  if (nrows == 0) {
    if (dot) B2S_CUDA(cudaMemsetAsync(dot_out, 0, vt == B2S_F32 ? 4 : 8, st));
    return B2S_OK;
  }
  if (plan == nullptr) {
    if (vt == B2S_F32) return dispatch_rowgroup<float>(it, pt, nrows, nnz, indptr, indices, vals, x, y, st);
    return dispatch_rowgroup<double>(it, pt, nrows, nnz, indptr, indices, vals, x, y, st);
  }
  const int64_t ntiles = b2s_spmv_plan_tiles(vt, nrows, nnz);
  const int vec_ok = (nnz == 0) || (aligned16(indices) && aligned16(vals));
  if (vt == B2S_F32) {
    if (dot) return dispatch_idx<float, true>(it, pt, g_cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, w, dot_out, ws, st);
    return dispatch_idx<float, false>(it, pt, g_cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, nullptr, nullptr, nullptr, st);
  }
  if (dot) return dispatch_idx<double, true>(it, pt, g_cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, w, dot_out, ws, st);
  return dispatch_idx<double, false>(it, pt, g_cfg, ntiles, indptr, indices, vals, x, y, plan, vec_ok, nullptr, nullptr, nullptr, st);
}

int b2s_spmv_csr(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                 const void* indices, const void* vals, const void* x, void* y, const int32_t* plan,
                 void* stream) {
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, nullptr, nullptr, plan, nullptr,
                   stream, false);
}

int b2s_spmv_csr_dot(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y, const void* w, void* dot_out,
                     const int32_t* plan, void* ws, void* stream) {
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, w, dot_out, plan, ws, stream, true);
}

}  // extern "C"
