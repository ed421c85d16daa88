// spmv_common.cuh -- declarations shared by the SpMV plan / C ABI (spmv.cu) and the per-value-type kernel
// translation units (spmv_f32.cu, spmv_f64.cu).
#pragma once
#include "common.cuh"
#include <limits.h>
#include <string.h>
#include <mutex>

namespace b2s {

constexpr int kMaxDevices = 64;
constexpr int kShortRowMax = 32;   // longest row the one-lane-per-row path takes

// pad: row-shape code of the tile written by the plan: L > 0 = every row has exactly L entries;
//      -M < 0 = rows differ but none is longer than M <= kShortRowMax (one-lane-per-row path); 0 = anything else.
struct __align__(16) PlanEntry {
  long long k;   // first nonzero of the tile's first row
  int row;       // first row of the tile
  int pad;
};

// ---------------------------------------------------------------------------------------------
// Tile configurations.  X(ID, KIND, A, B, C, D, XL)
//   KIND 0 (LDG) : A = THREADS, B = GROUPS (4-nnz groups per thread), C = MINB, D = SCALAR mapping flag
//                  CAP = 4*A*B
//   KIND 1 (TMA) : A = consumer warps, B = 16-byte groups per consumer thread, C = STAGES, D = MINB,
//                  XL = x-gather flavour (0 ld.global.nc, 1 .nc.L1::no_allocate, 2 .cg)
//                  CAP = (16/sizeof V) * 32*A * B
// T = CAP - 4 merge items per tile.  The four defaults are instantiated for every index type; the others
// (tuning sweeps, tools/, b2s_spmv_set_config) only for int32 indices/indptr.
// ---------------------------------------------------------------------------------------------
#define B2S_SPMV_CONFIGS(X) \
  X(0, 1, 4, 4, 2, 6, 0)    \
  X(1, 0, 128, 4, 6, 0, 0)  \
  X(2, 0, 128, 2, 8, 1, 0)  \
  X(3, 1, 4, 6, 2, 4, 0)    \
  X(4, 1, 4, 8, 2, 3, 0)    \
  X(5, 1, 4, 3, 2, 8, 0)    \
  X(6, 1, 8, 4, 2, 2, 0)    \
  X(7, 1, 4, 3, 2, 8, 2)    \
  X(8, 1, 4, 8, 2, 3, 2)    \
  X(9, 1, 4, 4, 2, 6, 2)    \
  X(10, 1, 4, 2, 3, 8, 0)   \
  X(11, 1, 8, 4, 2, 3, 0)   \
  X(12, 1, 8, 4, 2, 3, 2)   \
  X(13, 1, 8, 4, 2, 2, 2)
struct TileCfgRt { int kind, a, b, c, d, xl; };
static const TileCfgRt kCfgs[] = {
#define X(ID, K, A, B, C, D, XL) {K, A, B, C, D, XL},
    B2S_SPMV_CONFIGS(X)
#undef X
};
static constexpr int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);
static constexpr int kDefaultCfgF64 = 0;   // 4 consumer warps x 4 groups, 2 stages, 6 CTAs/SM (CAP 1024)
static constexpr int kDefaultCfgF32 = 5;   // 4 consumer warps x 3 groups, 2 stages, 8 CTAs/SM (CAP 1536)
// fp64 matrices of short rows with column locality (stencils, narrow bands: the one-lane-per-row path).  The shape "8
// consumer warps x 4 groups, 3 CTAs/SM" (cfg 11) is 1.5-3 % faster on the bare product (L5 129.4 vs 131.3 us, banded
// 11/row 240.1 vs 247.5 us on one box) but the fused-dot variant CG runs on lost 6 % with it (PDE4096 2167 vs
// 2301..2374 it/s), so the default shape stays.
static constexpr int kShortCfgF64 = kDefaultCfgF64;
// scattered matrices (plan statistic > 16 distinct x lines per warp gather) want many gathers in flight per thread:
static constexpr int kScatterCfgF64 = 2;   // LDG tiles, 128 threads x 8 nnz, scalar mapping
static constexpr int kScatterCfgF32 = 8;   // TMA tiles, 4 warps x 8 groups = 32 gathers/thread, x through ld.global.cg
                                           // (R32 fp32 on B200: .cg 1265 us, .nc 1355 us, .nc.L1::no_allocate 2744 us)
// scattered SHORT rows (one lane per row): x through ld.global.cg -- the L1-allocating path thrashes on random columns
// (column blocks of an R32 shard, 16 per row: 1194 us with .nc vs 706 us with .cg) -- and 8 consumer warps on tiles of
// 4096 (fp32) / 2048 (fp64) entries, 2 CTAs/SM: a tile's rows spread over 256 lanes, i.e. fewer dependent gather rounds
// per lane.  10M rows x 4 random entries (profiles/r02_sweep_r4.txt): fp32 172 us (cfg 13) vs 275 us (cfg 7: 4 warps x 3
// groups) vs 224 us without a plan; fp64 187 us (cfg 13) vs 293 us (cfg 9) vs 234 us.  Tiles of 8192 entries (8 warps x 8
// groups, 16 warps x 4 groups) were measured too: slower (181 / 183 us).
static constexpr int kScatterShortCfgF32 = 13;
static constexpr int kScatterShortCfgF64 = 13;

static inline int cfg_cap(int c, int vt) {
  const TileCfgRt& k = kCfgs[c];
  if (k.kind == 0) return 4 * k.a * k.b;
  return (vt == B2S_F32 ? 4 : 2) * 32 * k.a * k.b;
}
static inline int cfg_T(int c, int vt) { return cfg_cap(c, vt) - 4; }

// Tile processing order + the exchange fused into the TMA kernel (multi-GPU; csrc/peer.cu owns the buffers).
//  * tiles are visited range by range; ranges [0, n_free) only read locally valid x, the rest may read columns
//    that other GPUs push into this GPU's x buffer: before its first such tile each CTA's producer polls the
//    n_flags arrival flags (local memory, written remotely) until they reach this launch's epoch;
//  * send side: CTA b (< n_sends) copies send_count[b] elements from send_src[b] (local x) to send_dst[b]
//    (a peer-mapped pointer into the neighbour's x buffer) after the neighbour acknowledged the previous epoch
//    (send_ack[b], local), then stores the epoch to send_flag[b] (remote); CTA 0 stores epoch-1 to the n_acks
//    remote acknowledgement words of the GPUs that push into this one;
//  * epoch = *epoch_ctr + epoch_add (device counter; the last CTA stores the epoch back when epoch_bump) or
//    `expect` when epoch_ctr is NULL.  All zero = ordinary launch.
struct TileOrder {
  int nranges, n_free, n_flags, n_sends;
  int n_acks, accumulate, epoch_add, epoch_bump;
  int n_push, pad0;          // dedicated pusher CTAs (set by the launcher: min(n_sends, grid - 1); they take no tiles)
  long long lo[6], hi[6];
  const unsigned long long* flag[8];
  const void* send_src[4];
  void* send_dst[4];
  long long send_count[4];
  unsigned long long* send_flag[4];
  const unsigned long long* send_ack[4];
  unsigned long long* ack_out[8];
  unsigned long long* epoch_ctr;
  unsigned int* ticket;
  unsigned long long expect;
  unsigned long long* error;
};

struct SpmvArgs {
  int64_t ntiles, nrows, nnz;
  int64_t tile_lo, tile_hi;  // tile sub-range to run (TMA kernels); [0, ntiles) for a whole SpMV
  const TileOrder* order;    // optional explicit tile order + fused exchange (TMA kernels); NULL = [tile_lo, tile_hi)
  const void *indptr, *indices, *vals, *x;
  void* y;
  const PlanEntry* plan;
  int vec_ok;
  int flavor;      // TMA kernel flavour chosen by the plan: 0 generic, 1 uniform rows, 2 short rows
  int accumulate;  // y += A x (TMA kernels only)
  int waves;       // tuning hook: LDG kind grid-stride waves / TMA kind CTAs-per-SM cap (0 = automatic)
  const void* w;
  void* dot_out;
  void* ws;
  cudaStream_t st;
};

// per-value-type entry points (spmv_f32.cu / spmv_f64.cu)
int spmv_launch_f32(bool dot, int it, int pt, int cfg, const SpmvArgs& a);
int spmv_launch_f64(bool dot, int it, int pt, int cfg, const SpmvArgs& a);
int spmv_rowgroup_f32(int it, int pt, int64_t nrows, int64_t nnz, const void* indptr, const void* indices,
                      const void* vals, const void* x, void* y, cudaStream_t st);
int spmv_rowgroup_f64(int it, int pt, int64_t nrows, int64_t nnz, const void* indptr, const void* indices,
                      const void* vals, const void* x, void* y, cudaStream_t st);

}  // namespace b2s
