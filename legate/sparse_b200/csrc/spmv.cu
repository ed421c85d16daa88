// spmv.cu -- CSR SpMV for sm_100a: the plan (tile boundaries, row-shape codes, kernel choice) and the C ABI.
// The kernels live in spmv_kernels.cuh and are instantiated per value type in spmv_f32.cu / spmv_f64.cu.
//
// Replaces CSRSpMVRowSplit::gpu_variant (reference src/sparse/array/csr/spmv.cu:24-123, a cusparseSpMV call) and
// the partition helpers behind it (sparse/partition.py:56-208, src/sparse/partition/*.cu).
#include "spmv_common.cuh"
#include <atomic>
#include <stdlib.h>

namespace b2s {

// tuning hooks (tools/, tests): -1 / 0 = automatic
static std::atomic<int> g_cfg{-1};
static std::atomic<int> g_waves{0};

static inline int resolve_cfg(int vt, bool scattered, bool likely_short) {
  const int forced = g_cfg.load();
  if (forced >= 0) return forced;
  if (scattered && likely_short) return vt == B2S_F32 ? kScatterShortCfgF32 : kScatterShortCfgF64;
  if (scattered) return vt == B2S_F32 ? kScatterCfgF32 : kScatterCfgF64;
  if (likely_short && vt == B2S_F64) return kShortCfgF64;
  return vt == B2S_F32 ? kDefaultCfgF32 : kDefaultCfgF64;
}

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

// Second plan pass: the row-shape code of every tile (PlanEntry::pad).  L > 0: every row has the same length L
// (ELL-like tiles: fixed-degree graphs, interior rows of banded matrices) -- reducible straight from registers;
// -M: rows differ, none longer than M <= 32 -- one lane per row; 0: anything else.
// stats[0] += tiles the uniform register path can take, stats[1] += tiles the one-lane-per-row path can take.
template <typename P>
__global__ void spmv_plan_shape_kernel(const P* __restrict__ indptr, int64_t ntiles, PlanEntry* __restrict__ plan,
                                       int ept, unsigned long long* __restrict__ stats) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int r0 = plan[t].row, r1 = plan[t + 1].row;
  int code = 0;
  bool is_short = false;
  if (r1 > r0) {
    const long long len0 = (long long)indptr[r0 + 1] - (long long)indptr[r0];
    bool same = true;
    long long mx = len0;
    for (int r = r0 + 1; r < r1; r++) {
      const long long len = (long long)indptr[r + 1] - (long long)indptr[r];
      same = same && len == len0;
      mx = len > mx ? len : mx;
    }
    is_short = mx <= kShortRowMax;
    if (same && len0 > 0 && len0 < 32768) code = (int)len0;
    else if (is_short) code = -(int)(mx > 0 ? mx : 1);
  }
  plan[t].pad = code;
  // tiles the register fast path can take: L = ept * 2^s (s <= 5) and a 16-byte aligned first nonzero
  const int L = code > 0 ? code : 0;
  const int lpr = L / ept;
  if (L > 0 && lpr * ept == L && lpr <= 32 && (lpr & (lpr - 1)) == 0 && (plan[t].k & 3) == 0) atomicAdd(stats, 1ull);
  if (is_short) atomicAdd(stats + 1, 1ull);
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

// Column-locality statistic: mean number of distinct 128-byte lines of x touched by 32 consecutive
// nonzeros (one warp-wide gather), sampled at up to 4096 evenly spaced positions.  ~5 for stencils,
// 32 for uniformly random columns.  Scattered matrices are bound by the L1 tag stage (one lookup per
// distinct line), not by HBM: they get tile shapes with many gathers in flight per thread.
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
  int use_rowgroup; // 1: forced to the plan-free row-group kernel (tools / tests)
  int flavor;       // TMA kernel flavour: 0 generic, 1 uniform-row register path, 2 one-lane-per-row (short rows)
  int scattered;    // 1: > 16 distinct x lines per warp-wide gather
  int64_t nrows, ncols, nnz, ntiles;
  int64_t uniform_tiles, short_tiles;
  double lines_per_warp;
  const PlanEntry* dev;
  // row chunks for pipelined host<->device SpMV: chunk c = tiles [ctile[c], ctile[c+1]), rows [crow[c], crow[c+1]),
  // reading x only inside [ccol_lo[c], ccol_hi[c])
  int nchunks;
  int64_t ctile[kPlanChunks + 1], crow[kPlanChunks + 1], ccol_lo[kPlanChunks], ccol_hi[kPlanChunks];
};
static constexpr uint32_t kPlanMagic = 0xB2005A17u;

static int64_t tiles_for(int cfg, int vt, int64_t nrows, int64_t nnz) {
  if (nrows <= 0 || nnz < 0) return 0;
  const int64_t T = cfg_T(cfg, vt);
  return (nrows + nnz + T - 1) / T;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

// Debug/tuning hooks (not part of the documented ABI; used by tools/ sweep scripts).
// cfg < 0 restores automatic selection.
int b2s_spmv_set_config(int cfg, int waves) {
  if (cfg >= kNumCfgs) { set_error("config %d out of range [0,%d)", cfg, kNumCfgs); return B2S_EINVAL; }
  g_cfg.store(cfg < 0 ? -1 : cfg);
  g_waves.store(waves < 0 ? 0 : waves);
  return B2S_OK;
}
int b2s_spmv_get_config(void) { return g_cfg.load(); }
int b2s_spmv_num_configs(void) { return kNumCfgs; }

int64_t b2s_spmv_plan_tiles(int vt, int64_t nrows, int64_t nnz) {
  // upper bound over the configurations plan_create may pick (it decides after sampling the matrix)
  int64_t m = 0;
  const int forced = g_cfg.load();
  const int cands[6] = {forced >= 0 ? forced : (vt == B2S_F32 ? kDefaultCfgF32 : kDefaultCfgF64),
                        vt == B2S_F32 ? kScatterCfgF32 : kScatterCfgF64, vt == B2S_F32 ? kDefaultCfgF32 : kDefaultCfgF64,
                        kScatterCfgF32, vt == B2S_F32 ? kScatterShortCfgF32 : kScatterShortCfgF64,
                        vt == B2S_F32 ? kDefaultCfgF32 : kShortCfgF64};
  for (int c : cands) { const int64_t t = tiles_for(c, vt, nrows, nnz); m = t > m ? t : m; }
  return m;
}

int64_t b2s_spmv_plan_bytes(int vt, int64_t nrows, int64_t nnz) {
  // (ntiles + 1) entries + one 16-byte slot for the plan statistics
  return (b2s_spmv_plan_tiles(vt, nrows, nnz) + 2) * (int64_t)sizeof(PlanEntry);
}

static int plan_create_impl(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                            const void* indices, void* plan_buf, void* stream, void** plan_out, int flags) {
  B2S_CHECK_ARG(plan_out != nullptr, "plan_out is NULL");
  *plan_out = nullptr;
  if (int rc = check_common(vt, it, pt, nrows, ncols, nnz, indptr, indices, indices, plan_out, plan_out)) return rc;
  B2S_CHECK_ARG(plan_buf != nullptr && aligned16(plan_buf), "plan buffer must be non-NULL and 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  PlanEntry* dev = (PlanEntry*)plan_buf;
  // the statistics slot sits after the largest plan this buffer may hold
  unsigned long long* stat = (unsigned long long*)(dev + b2s_spmv_plan_tiles(vt, nrows, nnz) + 1);
  double lines = 0.0;
  int64_t uniform_tiles = 0, short_tiles = 0;
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
  // the deep-gather tile shapes are for scattered LONG rows; scattered short rows (e.g. one column block of a
  // column-blocked random shard, ~4 entries per row) take the one-lane-per-row path of the default shape
  const bool likely_short = nnz <= 20 * nrows;
  int cfg = resolve_cfg(vt, scattered, likely_short);
  // B2S_PLAN_TMA_ONLY: the caller needs the TMA tile kernel (y += A x, fused exchange): swap an LDG-kind choice
  // for the deep-gather TMA shape
  if ((flags & B2S_PLAN_TMA_ONLY) && kCfgs[cfg].kind != 1) cfg = kScatterCfgF32;
  const int64_t ntiles = tiles_for(cfg, vt, nrows, nnz);
  // 2. tile boundaries + row-shape codes for the chosen tile shape
  if (ntiles > 0) {
    const int64_t T = cfg_T(cfg, vt);
    const unsigned grid = (unsigned)((ntiles + 1 + 255) / 256);
    if (pt == B2S_I32) spmv_plan_kernel<int32_t><<<grid, 256, 0, st>>>(nrows, (const int32_t*)indptr, T, ntiles, dev);
    else               spmv_plan_kernel<int64_t><<<grid, 256, 0, st>>>(nrows, (const int64_t*)indptr, T, ntiles, dev);
    B2S_LAUNCH_CHECK();
    B2S_CUDA(cudaMemsetAsync(stat, 0, 16, st));
    const unsigned gu = (unsigned)((ntiles + 255) / 256);
    const int ept = vt == B2S_F32 ? 4 : 2;
    if (pt == B2S_I32) spmv_plan_shape_kernel<int32_t><<<gu, 256, 0, st>>>((const int32_t*)indptr, ntiles, dev, ept, stat);
    else               spmv_plan_shape_kernel<int64_t><<<gu, 256, 0, st>>>((const int64_t*)indptr, ntiles, dev, ept, stat);
    B2S_LAUNCH_CHECK();
    unsigned long long ut[2] = {0, 0};
    B2S_CUDA(cudaMemcpyAsync(ut, stat, sizeof(ut), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    uniform_tiles = (int64_t)ut[0];
    short_tiles = (int64_t)ut[1];
  }
  PlanHandle* h = new PlanHandle();
  h->magic = kPlanMagic;
  h->vt = vt; h->it = it; h->pt = pt; h->cfg = cfg;
  h->use_rowgroup = 0;
  h->scattered = scattered ? 1 : 0;
  // kernel flavour.  1: >= 25% of the tiles are ELL-like, or rows are long (>= 12) / of even mean length (the
  // generic reduce then wants the bank-skewed walk); 2: at least half of the tiles hold only short rows (<= 32
  // entries: stencils up to 27 points, narrow bands, column blocks of random shards) and the uniform path does
  // not already cover them;
  // 0: everything else (irregular long rows).
  const int64_t meanL = nrows > 0 ? (nnz + nrows / 2) / nrows : 0;
  int flavor = 0;
  if (ntiles > 0) {
    if (uniform_tiles * 4 >= ntiles) flavor = 1;
    else if (short_tiles * 2 >= ntiles) flavor = 2;
    else if (meanL >= 12 || (meanL > 0 && (meanL & 1) == 0)) flavor = 1;
  }
  h->flavor = flavor;
  h->nrows = nrows; h->ncols = ncols; h->nnz = nnz; h->ntiles = ntiles;
  h->uniform_tiles = uniform_tiles; h->short_tiles = short_tiles;
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

int b2s_spmv_plan_create(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                         const void* indices, void* plan_buf, void* stream, void** plan_out) {
  return plan_create_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, plan_buf, stream, plan_out, 0);
}

int b2s_spmv_plan_create_ex(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                            const void* indices, void* plan_buf, void* stream, void** plan_out, int flags) {
  return plan_create_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, plan_buf, stream, plan_out, flags);
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

/* out[0] = tile config, out[1] = bit0 row-group kernel forced, bit1 uniform-row flavour, bit2 scattered, bit3 short-row
 * flavour, bit4 TMA kind; out[2] = ntiles, out[3] = 1000 * mean distinct x lines per 32 consecutive nonzeros */
int b2s_spmv_plan_info(const void* plan, int64_t* out4_host) {
  const PlanHandle* h = (const PlanHandle*)plan;
  B2S_CHECK_ARG(h && h->magic == kPlanMagic && out4_host, "bad plan handle / out pointer");
  out4_host[0] = h->cfg;
  out4_host[1] = h->use_rowgroup + 2 * (h->flavor == 1) + 4 * h->scattered + 8 * (h->flavor == 2) + 16 * (kCfgs[h->cfg].kind == 1);
  out4_host[2] = h->ntiles;
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

/* force the TMA kernel flavour of a plan (tools / tests): 0 generic, 1 uniform, 2 short rows; < 0 leaves it */
int b2s_spmv_plan_set_flavor(void* plan, int flavor) {
  PlanHandle* h = (PlanHandle*)plan;
  B2S_CHECK_ARG(h && h->magic == kPlanMagic, "bad plan handle");
  B2S_CHECK_ARG(flavor <= 2, "flavor must be 0, 1 or 2");
  if (flavor >= 0) h->flavor = flavor;
  return B2S_OK;
}

static int spmv_impl(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y, const void* w, void* dot_out,
                     const void* plan, void* ws, void* stream, bool dot, int64_t tile_lo = 0, int64_t tile_hi = -1,
                     const TileOrder* order = nullptr, int accumulate = 0) {
  if (int rc = check_common(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dot) {
    B2S_CHECK_ARG(ws != nullptr && dot_out != nullptr, "ws/dot_out is NULL");
    B2S_CHECK_ARG(nrows == 0 || w != nullptr, "w is NULL");
    B2S_CHECK_ARG(!accumulate, "the fused inner product is not defined for y += A x");
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
  if (rowgroup && (order || accumulate)) {
    set_error("fused exchange / accumulate need the TMA tile kernel (a plan and 16-byte aligned arrays)");
    return B2S_EUNSUPPORTED;
  }
  if (rowgroup) {
    int rc = (vt == B2S_F32) ? spmv_rowgroup_f32(it, pt, nrows, nnz, indptr, indices, vals, x, y, st)
                             : spmv_rowgroup_f64(it, pt, nrows, nnz, indptr, indices, vals, x, y, st);
    if (rc || !dot) return rc;
    return b2s_dot(vt, nrows, w, y, dot_out, ws, stream);
  }
  SpmvArgs a;
  a.ntiles = h->ntiles;
  a.tile_lo = tile_lo;
  a.tile_hi = tile_hi < 0 ? h->ntiles : tile_hi;
  a.order = order;
  const bool tma = kCfgs[h->cfg].kind == 1;
  if ((order || accumulate) && !tma) {
    set_error("fused exchange / accumulate need a TMA tile plan (this plan uses LDG tile config %d)", h->cfg);
    return B2S_EUNSUPPORTED;
  }
  B2S_CHECK_ARG(a.tile_lo >= 0 && a.tile_lo <= a.tile_hi && a.tile_hi <= h->ntiles, "tile range out of bounds");
  B2S_CHECK_ARG((a.tile_lo == 0 && a.tile_hi == h->ntiles) || tma, "tile sub-ranges need a TMA tile plan");
  if (a.tile_lo == a.tile_hi && !order) return B2S_OK;
  a.nrows = nrows; a.nnz = nnz;
  a.indptr = indptr; a.indices = indices; a.vals = vals; a.x = x; a.y = y;
  a.plan = h->dev;
  a.vec_ok = aligned ? 1 : 0;
  a.flavor = h->flavor;
  a.accumulate = accumulate;
  a.waves = g_waves.load();
  a.w = dot ? w : nullptr; a.dot_out = dot ? dot_out : nullptr; a.ws = dot ? ws : nullptr;
  a.st = st;
  return vt == B2S_F32 ? spmv_launch_f32(dot, it, pt, h->cfg, a) : spmv_launch_f64(dot, it, pt, h->cfg, a);
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

/* SpMV with the x exchange fused into the kernel (see b2s_fuse_desc in include/b200sparse.h). */
int b2s_spmv_csr_fused(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                       const void* indices, const void* vals, const void* x, void* y, const void* w, void* dot_out,
                       const void* plan, void* ws, const b2s_fuse_desc* d, void* stream) {
  B2S_CHECK_ARG(plan != nullptr && d != nullptr, "b2s_spmv_csr_fused needs a plan and a descriptor");
  B2S_CHECK_ARG(d->nranges >= 1 && d->nranges <= 6 && d->n_free >= 0 && d->n_free <= d->nranges, "bad tile ranges");
  B2S_CHECK_ARG(d->n_flags >= 0 && d->n_flags <= 8 && d->n_sends >= 0 && d->n_sends <= 4 && d->n_acks >= 0 && d->n_acks <= 8,
                "too many flags / sends / acks");
  const bool exchanging = (d->n_flags | d->n_sends | d->n_acks) != 0;
  B2S_CHECK_ARG(!exchanging || d->error != nullptr, "exchange without an error word");
  B2S_CHECK_ARG(!exchanging || d->epoch_ctr != nullptr || d->expect > 0, "exchange without an epoch");
  B2S_CHECK_ARG(!(d->epoch_ctr && d->epoch_bump) || d->ticket != nullptr, "epoch_bump needs a ticket word");
  TileOrder o;
  memset(&o, 0, sizeof(o));
  o.nranges = d->nranges; o.n_free = d->n_free; o.n_flags = d->n_flags; o.n_sends = d->n_sends; o.n_acks = d->n_acks;
  o.epoch_add = d->epoch_add; o.epoch_bump = d->epoch_bump;
  for (int i = 0; i < d->nranges; i++) {
    o.lo[i] = d->ranges[2 * i]; o.hi[i] = d->ranges[2 * i + 1];
    B2S_CHECK_ARG(o.lo[i] >= 0 && o.hi[i] >= o.lo[i], "bad tile range %d", i);
  }
  for (int i = 0; i < d->n_flags; i++) { B2S_CHECK_ARG(d->flag[i], "NULL arrival flag %d", i); o.flag[i] = (const unsigned long long*)d->flag[i]; }
  for (int i = 0; i < d->n_sends; i++) {
    B2S_CHECK_ARG(d->send_src[i] && d->send_dst[i] && d->send_flag[i] && d->send_ack[i] && d->send_count[i] >= 0, "bad send %d", i);
    o.send_src[i] = d->send_src[i]; o.send_dst[i] = d->send_dst[i]; o.send_count[i] = d->send_count[i];
    o.send_flag[i] = (unsigned long long*)d->send_flag[i]; o.send_ack[i] = (const unsigned long long*)d->send_ack[i];
  }
  for (int i = 0; i < d->n_acks; i++) { B2S_CHECK_ARG(d->ack_out[i], "NULL ack word %d", i); o.ack_out[i] = (unsigned long long*)d->ack_out[i]; }
  o.epoch_ctr = (unsigned long long*)d->epoch_ctr;
  o.ticket = (unsigned int*)d->ticket;
  o.expect = d->expect;
  o.error = (unsigned long long*)d->error;
  const bool dot = w != nullptr || dot_out != nullptr;
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, w, dot_out, plan, ws, stream, dot, 0, -1,
                   &o, d->accumulate);
}

/* y += A x (TMA tile plans only): one column block of a column-blocked shard at a time */
int b2s_spmv_csr_add(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y, const void* plan, void* stream) {
  B2S_CHECK_ARG(plan != nullptr, "b2s_spmv_csr_add needs a plan");
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, nullptr, nullptr, plan, nullptr,
                   stream, false, 0, -1, nullptr, 1);
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
  // per-device copy streams and events, created once (under a lock: one host thread per GPU may call in)
  struct Pipe { cudaStream_t s_in = nullptr, s_out = nullptr; cudaEvent_t ev_in[kPlanChunks], ev_k[kPlanChunks], ev0; bool ok = false; };
  static Pipe pipes[kMaxDevices];
  static std::mutex pipes_mu;
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  B2S_CHECK_ARG(dev >= 0 && dev < kMaxDevices, "device ordinal out of range");
  Pipe& P = pipes[dev];
  {
    std::lock_guard<std::mutex> g(pipes_mu);
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
  }
  // B2S_PIPE_DIRECT=1: y straight into the caller's host buffer.  When y_host is page-locked and mapped (cudaHostAlloc /
  // cudaHostRegister; torch's pin_memory) the tiles of a stage store their rows over PCIe themselves instead of a device
  // buffer + a D2H copy per stage.  Measured (profiles/r02_e2e_direct.txt): NOT faster -- with the upstream direction busy
  // either way, a 5 MB H2D copy takes 137 us (36 GB/s) instead of 99 us, and the stores of a stage take as long; 2.32..2.47
  // ms per product vs 2.23 ms with copy-engine D2H.  Kept as an option (a box with a slower copy engine may differ).
  void* y_map = nullptr;
  {
    const char* e = getenv("B2S_PIPE_DIRECT");
    const bool want = e && e[0] == '1';
    if (want && nrows > 0) {
      if (cudaHostGetDevicePointer(&y_map, y_host, 0) != cudaSuccess) { y_map = nullptr; cudaGetLastError(); }
    }
  }
  const bool direct = y_map != nullptr;
  // B2S_PIPE_TRACE=1: print when each chunk's H2D / tiles / D2H started and ended (debugging the overlap)
  const char* trace_env = getenv("B2S_PIPE_TRACE");
  const bool trace = trace_env != nullptr && trace_env[0] == '1';
  cudaEvent_t tr[kPlanChunks][6];
  cudaEvent_t tr0 = nullptr;
  if (trace) {
    B2S_CUDA(cudaEventCreate(&tr0));
    for (int c = 0; c < h->nchunks; c++) for (int q = 0; q < 6; q++) B2S_CUDA(cudaEventCreate(&tr[c][q]));
    B2S_CUDA(cudaEventRecord(tr0, st));
  }
  // the copy streams start after everything already queued on the compute stream
  B2S_CUDA(cudaEventRecord(P.ev0, st));
  B2S_CUDA(cudaStreamWaitEvent(P.s_in, P.ev0, 0));
  B2S_CUDA(cudaStreamWaitEvent(P.s_out, P.ev0, 0));
  // pipeline schedule over the plan's 16 chunks.  Small copies run well below the duplex PCIe rate (5 MB stages: ~75
  // GB/s aggregate, 256 MB copies: 99 GB/s) but long first / last stages leave one direction idle while the pipeline
  // fills and drains, so the default is GRADED: 1, 2, 2, 2, 2, 2, 2, 2, 1 chunks per stage (short first and last stages:
  // the pipeline fills and drains in 1/16 of the vector; 10 MB ones in between).  Copies alone, staged the same way,
  // need 1.86 ms on the bench box (tools/pcie_duplex_chunks.py), this pipeline 2.02 ms.  B2S_PIPE_CHUNKS = n (1..16)
  // asks for n equal stages instead, B2S_PIPE_PATTERN = "a,b,..." (sum 16) for any other grading.
  int bounds[kPlanChunks + 2];
  int nst = 0;
  bounds[0] = 0;
  {
    int stages = 0;
    if (const char* e = getenv("B2S_PIPE_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= kPlanChunks) stages = v; }
    if (stages <= 0 && direct) stages = 6;          // measured best for the direct-store option
    int pattern[kPlanChunks], npat = 0, psum = 0;   // B2S_PIPE_PATTERN = "1,3,4,4,3,1": chunks per stage (sum 16)
    if (const char* e = getenv("B2S_PIPE_PATTERN")) {
      const char* q = e;
      while (*q && npat < kPlanChunks) {
        const int v = atoi(q);
        if (v <= 0) { npat = 0; break; }
        pattern[npat++] = v; psum += v;
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
      }
      if (psum != h->nchunks) npat = 0;
    }
    if (npat > 0) {
      int c = 0;
      for (int i = 0; i < npat; i++) { c += pattern[i]; bounds[++nst] = c; }
    } else if (stages > 0 || h->nchunks != kPlanChunks) {
      if (stages <= 0) stages = h->nchunks;
      const int grp = (h->nchunks + stages - 1) / stages;
      for (int c = grp; c < h->nchunks; c += grp) bounds[++nst] = c;
      bounds[++nst] = h->nchunks;
    } else {
      static const int graded[9] = {1, 2, 2, 2, 2, 2, 2, 2, 1};
      int c = 0;
      for (int i = 0; i < 9; i++) { c += graded[i]; bounds[++nst] = c; }
    }
  }
  // Copy boundaries sit on multiples of 4 KB (x: rounded up, y: rounded down; the rows a stage finishes past its aligned
  // end travel with the next stage): D2H copies that start inside a page run at ~40 GB/s next to a busy H2D direction,
  // page-aligned ones at ~47 (L5 product 2.20 -> 2.03 ms, profiles/r02_e2e_patterns.txt).  B2S_PIPE_ALIGN = n elements
  // overrides (0 = exact chunk boundaries).
  int64_t align = (int64_t)(4096 / sv);
  if (const char* e = getenv("B2S_PIPE_ALIGN")) { align = atoll(e); if (align < 0) align = 0; }
  int64_t copied = 0, y_done = 0;
  for (int sidx = 0; sidx < nst; sidx++) {
    const int c = bounds[sidx], ce = bounds[sidx + 1];            // chunks [c, ce)
    int64_t need = 0;
    for (int q = c; q < ce; q++) need = h->ccol_hi[q] > need ? h->ccol_hi[q] : need;
    if (align > 1) { need = (need + align - 1) / align * align; if (need > ncols) need = ncols; }
    if (need > copied) {
      if (trace) B2S_CUDA(cudaEventRecord(tr[c][0], P.s_in));
      B2S_CUDA(cudaMemcpyAsync((char*)x_dev + sv * copied, (const char*)x_host + sv * copied, sv * (size_t)(need - copied),
                               cudaMemcpyHostToDevice, P.s_in));
      if (trace) B2S_CUDA(cudaEventRecord(tr[c][1], P.s_in));
      copied = need;
      B2S_CUDA(cudaEventRecord(P.ev_in[c], P.s_in));
      B2S_CUDA(cudaStreamWaitEvent(st, P.ev_in[c], 0));
    }
    if (trace) B2S_CUDA(cudaEventRecord(tr[c][2], st));
    if (int rc = b2s_spmv_csr_tiles(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x_dev, direct ? y_map : y_dev,
                                    plan, h->ctile[c], h->ctile[ce], stream)) return rc;
    if (trace) B2S_CUDA(cudaEventRecord(tr[c][3], st));
    int64_t r0 = y_done, r1 = h->crow[ce];
    if (align > 1 && sidx + 1 < nst) r1 = r1 / align * align;
    if (r1 > r0 && !direct) {
      y_done = r1;
      B2S_CUDA(cudaEventRecord(P.ev_k[c], st));
      B2S_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_k[c], 0));
      if (trace) B2S_CUDA(cudaEventRecord(tr[c][4], P.s_out));
      B2S_CUDA(cudaMemcpyAsync((char*)y_host + sv * r0, (const char*)y_dev + sv * r0, sv * (size_t)(r1 - r0),
                               cudaMemcpyDeviceToHost, P.s_out));
      if (trace) B2S_CUDA(cudaEventRecord(tr[c][5], P.s_out));
    }
  }
  B2S_CUDA(cudaStreamSynchronize(P.s_out));
  B2S_CUDA(cudaStreamSynchronize(P.s_in));
  B2S_CUDA(cudaStreamSynchronize(st));
  if (trace) {
    for (int c = 0; c < h->nchunks; c++) {
      float t[6] = {-1, -1, -1, -1, -1, -1};
      for (int q = 0; q < 6; q++) if (cudaEventQuery(tr[c][q]) == cudaSuccess) cudaEventElapsedTime(&t[q], tr0, tr[c][q]);
      fprintf(stderr, "[b2s pipe] chunk %2d  h2d %7.3f..%7.3f  tiles %7.3f..%7.3f  d2h %7.3f..%7.3f ms\n", c, t[0], t[1], t[2],
              t[3], t[4], t[5]);
      for (int q = 0; q < 6; q++) cudaEventDestroy(tr[c][q]);
    }
    cudaEventDestroy(tr0);
    cudaGetLastError();   // events never recorded report an error on query: clear it
  }
  return B2S_OK;
}

int b2s_spmv_csr_dot(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y, const void* w, void* dot_out,
                     const void* plan, void* ws, void* stream) {
  return spmv_impl(vt, it, pt, nrows, ncols, nnz, indptr, indices, vals, x, y, w, dot_out, plan, ws, stream, true);
}

}  // extern "C"
