// gather_bench.cu -- what does a B200 allow for the irregular half of CSR SpMV?
//
// BASELINE config 4 (R32: 10M x 10M, 32 uniformly random columns per row, fp32) performs 320 M independent
// 4-byte reads of a 40 MB vector per product.  This standalone tool (no library code) measures the ceiling
// of exactly that access pattern, so the SpMV kernel's distance from it can be stated in numbers:
//
//   hash   : column = hash(global index) computed in registers -- nothing else is read: the pure gather rate
//   idx    : columns streamed from a 1.28 GB int32 array (coalesced 16-byte loads) -- gather + index stream
//   spmv   : columns AND values streamed, products summed per 32-entry row, y written -- the whole R32 SpMV
//            traffic with a trivial (8 lanes per row) reduction
//
// for several load flavours (ld.global.nc / .nc.L1::no_allocate / .cg / .ca), gathers in flight per thread (U),
// CTAs per SM, element widths (4 B / 8 B) and vector lengths (L2-resident 5/20/40 MB, and 80/160 MB which are
// not).  One line per case:  name  n_gathers  us  Ggather/s  sector-GB/s (32 B per gather).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/gather_bench tools/gather_bench.cu
//   tools/bin/gather_bench [filter-substring]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t v) {
  v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
  return v;
}

template <typename V, int FL> __device__ __forceinline__ V ldx(const V* p);
template <> __device__ __forceinline__ float ldx<float, 0>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ double ldx<double, 0>(const double* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldx<float, 1>(const float* p) {
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
template <> __device__ __forceinline__ double ldx<double, 1>(const double* p) {
  double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
template <> __device__ __forceinline__ float ldx<float, 2>(const float* p) { return __ldcg(p); }
template <> __device__ __forceinline__ double ldx<double, 2>(const double* p) { return __ldcg(p); }
template <> __device__ __forceinline__ float ldx<float, 3>(const float* p) { return __ldca(p); }
template <> __device__ __forceinline__ double ldx<double, 3>(const double* p) { return __ldca(p); }
template <> __device__ __forceinline__ float ldx<float, 4>(const float* p) {
  float v; asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
template <> __device__ __forceinline__ double ldx<double, 4>(const double* p) {
  double v; asm volatile("ld.global.nc.L1::evict_first.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}

// ---- pure gather: columns from a hash, U independent loads in flight per thread -------------------------
template <typename V, int FL, int U>
__global__ void __launch_bounds__(256) gather_hash(const V* __restrict__ x, uint32_t n, long long total, V* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  V acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride * U) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t h = hash32((uint32_t)(i + u * stride));
      const uint32_t c = (uint32_t)(((unsigned long long)h * n) >> 32);
      v[u] = ldx<V, FL>(x + c);
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u];
  }
  if (acc == (V)123456789) out[0] = acc;
}

// ---- gather with the index stream (int4 loads: 4 columns per lane per load, G loads in flight) -----------
template <typename V, int FL, int G>
__global__ void __launch_bounds__(256) gather_idx(const V* __restrict__ x, const int4* __restrict__ idx4, long long n4,
                                                  V* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  V acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * G) {
    int4 c[G];
#pragma unroll
    for (int g = 0; g < G; g++) c[g] = (i + g * stride < n4) ? __ldcs(idx4 + i + g * stride) : make_int4(0, 0, 0, 0);
    V v[G][4];
#pragma unroll
    for (int g = 0; g < G; g++) {
      v[g][0] = ldx<V, FL>(x + c[g].x); v[g][1] = ldx<V, FL>(x + c[g].y);
      v[g][2] = ldx<V, FL>(x + c[g].z); v[g][3] = ldx<V, FL>(x + c[g].w);
    }
#pragma unroll
    for (int g = 0; g < G; g++) acc += (v[g][0] + v[g][1]) + (v[g][2] + v[g][3]);
  }
  if (acc == (V)123456789) out[0] = acc;
}

// ---- the whole R32 traffic: columns + values streamed, 8 lanes x 4 entries per 32-entry row, y written -----
template <typename V> struct Vec4;
template <> struct Vec4<float> { typedef float4 T; };
template <> struct Vec4<double> { typedef double4 T; };
__device__ __forceinline__ float4 ld4(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ double4 ld4(const double4* p) {
  double2 a = __ldcs(reinterpret_cast<const double2*>(p)), b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
template <typename V, int FL, int G>
__global__ void __launch_bounds__(256) spmv_ell32(const V* __restrict__ x, const int4* __restrict__ idx4,
                                                  const typename Vec4<V>::T* __restrict__ val4, long long n4, V* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;   // multiple of 8: a row = 8 consecutive lanes
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * G) {
    int4 c[G];
    typename Vec4<V>::T a[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      const bool in = i + g * stride < n4;
      c[g] = in ? __ldcs(idx4 + i + g * stride) : make_int4(0, 0, 0, 0);
      if (in) a[g] = ld4(val4 + i + g * stride); else { a[g].x = 0; a[g].y = 0; a[g].z = 0; a[g].w = 0; }
    }
    V s[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      const V x0 = ldx<V, FL>(x + c[g].x), x1 = ldx<V, FL>(x + c[g].y), x2 = ldx<V, FL>(x + c[g].z), x3 = ldx<V, FL>(x + c[g].w);
      s[g] = (a[g].x * x0 + a[g].y * x1) + (a[g].z * x2 + a[g].w * x3);
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      V v = s[g];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      const long long j = i + g * stride;
      if ((threadIdx.x & 7) == 0 && j < n4) y[j >> 3] = v;
    }
  }
}

__global__ void fill_idx(int* idx, long long n, uint32_t ncols, uint32_t seed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t h = hash32((uint32_t)i * 2654435761U + seed);
    idx[i] = (int)(((unsigned long long)h * ncols) >> 32);
  }
}
template <typename V> __global__ void fill_val(V* v, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = (V)((i % 1000) * 1e-3);
}

static const char* g_filter = nullptr;
static int g_sms = 148;

template <typename F>
static void timeit(const char* name, long long gathers, double extra_bytes, F launch) {
  if (g_filter && !strstr(name, g_filter)) return;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  std::vector<float> ts;
  for (int i = 0; i < 7; i++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    ts.push_back(ms);
  }
  CK(cudaGetLastError());
  std::sort(ts.begin(), ts.end());
  const double t = ts[ts.size() / 2] * 1e-3;
  printf("%-46s gathers %11lld  med %9.1f us  min %9.1f us  %7.1f Ggather/s  sectors %8.1f GB/s  stream %7.1f GB/s\n",
         name, gathers, t * 1e6, ts[0] * 1e3, gathers / t / 1e9, gathers * 32.0 / t / 1e9, extra_bytes / t / 1e9);
  fflush(stdout);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

template <typename V, int FL, int U>
static void run_hash(const char* tag, const V* x, uint32_t n, long long total, V* out, int cps) {
  char name[128];
  snprintf(name, sizeof name, "hash %s n=%uM fl%d U%d cta/sm%d", tag, n / 1000000, FL, U, cps);
  timeit(name, total, 0.0, [&] { gather_hash<V, FL, U><<<g_sms * cps, 256>>>(x, n, total, out); });
}
template <typename V, int FL, int G>
static void run_idx(const char* tag, const V* x, uint32_t n, const int* idx, long long total, V* out, int cps) {
  char name[128];
  snprintf(name, sizeof name, "idx  %s n=%uM fl%d G%d(x4) cta/sm%d", tag, n / 1000000, FL, G, cps);
  timeit(name, total, total * 4.0, [&] { gather_idx<V, FL, G><<<g_sms * cps, 256>>>(x, (const int4*)idx, total / 4, out); });
}
template <typename V, int FL, int G>
static void run_spmv(const char* tag, const V* x, uint32_t n, const int* idx, const V* val, long long total, V* y, int cps) {
  char name[128];
  snprintf(name, sizeof name, "spmv %s n=%uM fl%d G%d(x4) cta/sm%d", tag, n / 1000000, FL, G, cps);
  timeit(name, total, total * (4.0 + sizeof(V)) + total / 32 * sizeof(V),
         [&] { spmv_ell32<V, FL, G><<<g_sms * cps, 256>>>(x, (const int4*)idx, (const typename Vec4<V>::T*)val, total / 4, y); });
}

template <typename V>
static void suite(const char* tag) {
  const long long total = 320000000LL;  // R32: 10M rows x 32
  int* idx = nullptr; V* val = nullptr; V* x = nullptr; V* y = nullptr;
  const uint32_t nmax = 40000000u;
  CK(cudaMalloc(&idx, total * sizeof(int)));
  CK(cudaMalloc(&val, total * sizeof(V)));
  CK(cudaMalloc(&x, (size_t)nmax * sizeof(V)));
  CK(cudaMalloc(&y, (size_t)(total / 32 + 8) * sizeof(V)));
  fill_val<V><<<g_sms * 8, 256>>>(val, total);
  fill_val<V><<<g_sms * 8, 256>>>(x, nmax);
  CK(cudaDeviceSynchronize());
  // (1) pure gather ceiling vs vector length (L2 residency): 1.25M .. 40M elements
  const uint32_t ns[] = {1250000u, 5000000u, 10000000u, 20000000u, 40000000u};
  for (uint32_t n : ns) {
    run_hash<V, 0, 16>(tag, x, n, total, y, 8);
  }
  // (2) load flavour / depth / occupancy at the R32 vector length
  const uint32_t n = 10000000u;
  run_hash<V, 0, 8>(tag, x, n, total, y, 8);
  run_hash<V, 0, 32>(tag, x, n, total, y, 8);
  run_hash<V, 0, 16>(tag, x, n, total, y, 4);
  run_hash<V, 0, 16>(tag, x, n, total, y, 2);
  run_hash<V, 1, 16>(tag, x, n, total, y, 8);
  run_hash<V, 2, 16>(tag, x, n, total, y, 8);
  run_hash<V, 3, 16>(tag, x, n, total, y, 8);
  run_hash<V, 4, 16>(tag, x, n, total, y, 8);
  // (3) with the column stream
  fill_idx<<<g_sms * 8, 256>>>(idx, total, n, 17u);
  CK(cudaDeviceSynchronize());
  run_idx<V, 0, 2>(tag, x, n, idx, total, y, 8);
  run_idx<V, 0, 4>(tag, x, n, idx, total, y, 8);
  run_idx<V, 0, 8>(tag, x, n, idx, total, y, 4);
  run_idx<V, 1, 4>(tag, x, n, idx, total, y, 8);
  run_idx<V, 2, 4>(tag, x, n, idx, total, y, 8);
  // (4) the whole R32 SpMV traffic
  run_spmv<V, 0, 2>(tag, x, n, idx, val, total, y, 8);
  run_spmv<V, 0, 4>(tag, x, n, idx, val, total, y, 8);
  run_spmv<V, 0, 4>(tag, x, n, idx, val, total, y, 4);
  run_spmv<V, 0, 8>(tag, x, n, idx, val, total, y, 4);
  run_spmv<V, 1, 4>(tag, x, n, idx, val, total, y, 8);
  run_spmv<V, 2, 4>(tag, x, n, idx, val, total, y, 8);
  // strong-scaling shard shape: 1/8 of the rows, full-length x
  run_spmv<V, 0, 4>(tag, x, n, idx, val, total / 8, y, 8);
  // (5) window-restricted columns (x working set 256 KB per CTA neighbourhood): the "banded random" contrast
  CK(cudaFree(idx)); CK(cudaFree(val)); CK(cudaFree(x)); CK(cudaFree(y));
}

int main(int argc, char** argv) {
  if (argc > 1) g_filter = argv[1];
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  g_sms = p.multiProcessorCount;
  printf("# %s, %d SMs, L2 %d MB, clock %d MHz\n", p.name, g_sms, p.l2CacheSize >> 20, p.clockRate / 1000);
  suite<float>("f32");
  suite<double>("f64");
  return 0;
}
