// sharded_spmv_test.cu -- drives the row-sharded SpMV + CG scalars through the C ABI ALONE: no Python, no
// torch.distributed.  One host thread per GPU rank (the library's threading contract), NCCL communicator owned by
// libb200sparse.so (b2s_comm_*), x all-gathered with b2s_allgather_x, shard products with b2s_spmv_csr, p.q
// all-reduced with b2s_allreduce_scalars.  Checked against a host loop (the oracle's arithmetic, spmv.cc:36-44).
//
//   nvcc -O2 -o sharded_spmv_test sharded_spmv_test.cu -I../../include -L../../legate/sparse_b200 -lb200sparse
//   ./sharded_spmv_test <nranks>      (nranks <= visible GPUs; 1 works on a single-GPU box)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include "b200sparse.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)
#define B2(x) do { int r_ = (x); if (r_ != B2S_OK) { fprintf(stderr, "%s:%d rc=%d %s\n", __FILE__, __LINE__, r_, b2s_last_error()); exit(3); } } while (0)

static const int64_t N = 200000;   // global rows; banded, offsets -301,-1,0,1,301 like tests/dist_worker.py
static const int OFFS[5] = {-301, -1, 0, 1, 301};
static const double VALS[5] = {1.0, -2.0, 3.0, -4.0, 5.0};

struct Shard { std::vector<int32_t> indptr, indices; std::vector<double> vals; int64_t lo, hi; };

static Shard make_shard(int rank, int nranks) {
  Shard s;
  const int64_t T = (N + nranks - 1) / nranks;
  s.lo = rank * T; s.hi = s.lo + T < N ? s.lo + T : N;
  s.indptr.push_back(0);
  for (int64_t r = s.lo; r < s.hi; r++) {
    for (int k = 0; k < 5; k++) {
      const int64_t c = r + OFFS[k];
      if (c >= 0 && c < N) { s.indices.push_back((int32_t)c); s.vals.push_back(VALS[k]); }
    }
    s.indptr.push_back((int32_t)s.indices.size());
  }
  return s;
}

static double xval(int64_t i) { return sin(0.001 * (double)i) + 0.5; }

static void rank_main(int rank, int nranks, const char* id, double* pq_out, double* err_out) {
  CK(cudaSetDevice(rank));
  void* comm = nullptr;
  B2(b2s_comm_init(rank, nranks, id, &comm));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  Shard s = make_shard(rank, nranks);
  const int64_t T = (N + nranks - 1) / nranks, nloc = s.hi - s.lo, nnz = (int64_t)s.indices.size();
  const int64_t npad = T * nranks;
  int32_t *d_ptr, *d_idx; double *d_val, *d_x, *d_y, *d_pq; void* d_ws; void* d_plan;
  CK(cudaMalloc(&d_ptr, sizeof(int32_t) * (nloc + 1)));
  CK(cudaMalloc(&d_idx, sizeof(int32_t) * nnz));
  CK(cudaMalloc(&d_val, sizeof(double) * nnz));
  CK(cudaMalloc(&d_x, sizeof(double) * npad));
  CK(cudaMalloc(&d_y, sizeof(double) * nloc));
  CK(cudaMalloc(&d_pq, sizeof(double)));
  CK(cudaMalloc(&d_ws, (size_t)b2s_ws_bytes()));
  CK(cudaMemset(d_ws, 0, (size_t)b2s_ws_bytes()));
  CK(cudaMemset(d_x, 0, sizeof(double) * npad));
  CK(cudaMalloc(&d_plan, (size_t)b2s_spmv_plan_bytes(B2S_F64, nloc, nnz)));
  CK(cudaMemcpy(d_ptr, s.indptr.data(), sizeof(int32_t) * (nloc + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, s.indices.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_val, s.vals.data(), sizeof(double) * nnz, cudaMemcpyHostToDevice));
  std::vector<double> xl(nloc);
  for (int64_t i = 0; i < nloc; i++) xl[i] = xval(s.lo + i);
  CK(cudaMemcpy(d_x + rank * T, xl.data(), sizeof(double) * nloc, cudaMemcpyHostToDevice));
  void* plan = nullptr;
  B2(b2s_spmv_plan_create(B2S_F64, B2S_I32, B2S_I32, nloc, N, nnz, d_ptr, d_idx, d_plan, st, &plan));
  // the sharded step: exchange, product fused with the local p.q, scalar all-reduce
  B2(b2s_allgather_x(comm, B2S_F64, d_x + rank * T, T, d_x, st));
  B2(b2s_spmv_csr_dot(B2S_F64, B2S_I32, B2S_I32, nloc, N, nnz, d_ptr, d_idx, d_val, d_x, d_y, d_x + s.lo, d_pq, plan, d_ws, st));
  B2(b2s_allreduce_scalars(comm, d_pq, 1, st));
  CK(cudaStreamSynchronize(st));
  std::vector<double> y(nloc);
  double pq = 0.0;
  CK(cudaMemcpy(y.data(), d_y, sizeof(double) * nloc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&pq, d_pq, sizeof(double), cudaMemcpyDeviceToHost));
  double err = 0.0;
  for (int64_t r = s.lo; r < s.hi; r++) {
    double ref = 0.0;
    for (int k = 0; k < 5; k++) { const int64_t c = r + OFFS[k]; if (c >= 0 && c < N) ref += VALS[k] * xval(c); }
    err = fmax(err, fabs(ref - y[r - s.lo]));
  }
  pq_out[rank] = pq;
  err_out[rank] = err;
  B2(b2s_spmv_plan_destroy(plan));
  B2(b2s_comm_destroy(comm));
  CK(cudaFree(d_ptr)); CK(cudaFree(d_idx)); CK(cudaFree(d_val)); CK(cudaFree(d_x)); CK(cudaFree(d_y));
  CK(cudaFree(d_pq)); CK(cudaFree(d_ws)); CK(cudaFree(d_plan));
}

int main(int argc, char** argv) {
  const int nranks = argc > 1 ? atoi(argv[1]) : 1;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (nranks < 1 || nranks > ndev) { fprintf(stderr, "need %d GPUs, have %d\n", nranks, ndev); return 77; }
  if (b2s_comm_nccl_version() == 0) { fprintf(stderr, "NCCL not loadable: %s\n", b2s_last_error()); return 4; }
  char id[128];
  B2(b2s_comm_unique_id(id));
  std::vector<double> pq(nranks), err(nranks);
  std::vector<std::thread> th;
  for (int r = 0; r < nranks; r++) th.emplace_back(rank_main, r, nranks, id, pq.data(), err.data());
  for (auto& t : th) t.join();
  // host reference of x . (A x)
  double ref = 0.0;
  for (int64_t r = 0; r < N; r++) {
    double yr = 0.0;
    for (int k = 0; k < 5; k++) { const int64_t c = r + OFFS[k]; if (c >= 0 && c < N) yr += VALS[k] * xval(c); }
    ref += xval(r) * yr;
  }
  double maxerr = 0.0;
  for (int r = 0; r < nranks; r++) {
    maxerr = fmax(maxerr, err[r]);
    if (fabs(pq[r] - ref) > 1e-9 * fabs(ref)) { fprintf(stderr, "rank %d: p.q %.17g vs %.17g\n", r, pq[r], ref); return 5; }
    if (pq[r] != pq[0]) { fprintf(stderr, "all-reduced scalar differs between ranks\n"); return 6; }
  }
  if (maxerr > 1e-11) { fprintf(stderr, "SpMV max err %.3e\n", maxerr); return 7; }
  printf("SHARDED_C_ABI_OK nranks=%d nccl=%d maxerr=%.2e pq=%.12g\n", nranks, b2s_comm_nccl_version(), maxerr, ref);
  return 0;
}
