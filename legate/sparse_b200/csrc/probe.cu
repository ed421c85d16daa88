// probe.cu -- measurement probe: the gather rate a B200 sustains on the access pattern of a scattered SpMV.
//
// bench.py times this next to the R32 product (BASELINE config 4: 320 M independent 4-byte reads of a 40 MB
// vector per SpMV) so the kernel's distance from what the memory system allows for that pattern -- not only from
// the HBM streaming roofline, which uniformly random columns cannot reach -- is measured in the same run.
// Columns are hashed from the global index in registers (nothing else is read), 16 independent loads in flight
// per thread, 8 CTAs of 256 threads per SM.  tools/gather_bench.cu is the standalone sweep this was distilled from.
#include "common.cuh"

namespace b2s {

__device__ __forceinline__ uint32_t probe_hash32(uint32_t v) {
  v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
  return v;
}

template <typename V, int U>
__global__ void __launch_bounds__(256) probe_gather_kernel(const V* __restrict__ x, uint32_t n, long long total, V* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  V acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride * U) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t h = probe_hash32((uint32_t)(i + u * stride));
      v[u] = __ldg(x + (uint32_t)(((unsigned long long)h * n) >> 32));
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u];
  }
  if (acc == (V)123456789) out[0] = acc;   // never true for the probe's inputs; keeps the loads alive
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_probe_gather(int vt, int64_t ncols, int64_t ngathers, const void* x_dev, void* out_dev, void* stream) {
  B2S_CHECK_ARG(vt == B2S_F32 || vt == B2S_F64, "bad value type code %d", vt);
  B2S_CHECK_ARG(ncols > 0 && ncols < (1LL << 32) && ngathers > 0 && x_dev && out_dev, "bad probe arguments");
  DeviceProps pr;
  if (int rc = get_props(&pr)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)pr.sm_count * 8u;
  if (vt == B2S_F32) probe_gather_kernel<float, 16><<<grid, 256, 0, st>>>((const float*)x_dev, (uint32_t)ncols, ngathers, (float*)out_dev);
  else               probe_gather_kernel<double, 16><<<grid, 256, 0, st>>>((const double*)x_dev, (uint32_t)ncols, ngathers, (double*)out_dev);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// cudaLimitMaxL2FetchGranularity of the current device: the size (32 / 64 / 128 bytes) L2 fetches from HBM on a miss.
// A hint to the driver; matters only for scattered reads of a vector larger than L2 (weak-scaled R32: x = 320 MB), where
// 128-byte fills move 4x the sectors the product consumes.  set_bytes = 0 only reads the limit back.
extern "C" int b2s_device_l2_fetch_granularity(int set_bytes, int64_t* current) {
  B2S_CHECK_ARG(set_bytes == 0 || set_bytes == 32 || set_bytes == 64 || set_bytes == 128, "granularity must be 0/32/64/128");
  if (set_bytes) B2S_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)set_bytes));
  size_t v = 0;
  B2S_CUDA(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
  if (current) *current = (int64_t)v;
  return B2S_OK;
}
