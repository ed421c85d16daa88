/*
 * oracle.c -- CPU restatement of the legate.sparse hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA path.  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so.  The product library
 * (libb200sparse.so) never links or calls anything in here.
 *
 * Each function restates one reference leaf task (file:line relative to the
 * reference checkout).  The reference itself cannot be compiled here (every
 * native file includes legate.h; legate.core/Legion/cuNumeric are absent), so
 * this restatement is pinned the way the reference's own tests pin the path:
 * against scipy.sparse on the reference's .mtx fixtures and seeded matrices
 * (tests/integration/test_csr_dot.py:25-45, test_csr_spgemm.py:24-32,
 * test_cg_solve.py:23-106) -- see tests/test_oracle_pinning.py and
 * tests/golden/make_golden.py.  Parity status: PINNED to scipy 1.18.1 golden
 * vectors + the test.mtx known answer; the reference binary itself was not run.
 *
 * Index/value dispatch: vt 0=f32 1=f64; it/pt 0=int32 1=int64 (columns / indptr).
 * CSR layout is plain scipy-style indptr (nrows+1) instead of the reference's
 * Rect<1> pos {lo,hi inclusive}: pos[i] == {indptr[i], indptr[i+1]-1}
 * (sparse/base.py:30-48, sparse/csr.py:426-440).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_EINVAL 1
#define ORC_ENOMEM 2

static inline int64_t ld_idx(const void* p, int wide, int64_t i)
{
  return wide ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i];
}
static inline void st_idx(void* p, int wide, int64_t i, int64_t v)
{
  if (wide) ((int64_t*)p)[i] = v; else ((int32_t*)p)[i] = (int32_t)v;
}

int orc_version(void) { return 1; }

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Copy `nbytes` from src to dst page by page inside a static OpenMP loop, so that a freshly allocated dst is
 * FIRST TOUCHED by all threads (pages spread over the NUMA nodes of the box instead of all landing on the node of
 * the one Python thread that filled the source).  Measurement hygiene for bench.py's CPU arms; no reference
 * counterpart (Legion places instances per processor). */
#include <string.h>
void orc_parallel_copy(void* dst, const void* src, int64_t nbytes)
{
  const int64_t page = 1 << 16;
  const int64_t npages = (nbytes + page - 1) / page;
  _Pragma("omp parallel for schedule(static)")
  for (int64_t p = 0; p < npages; p++) {
    const int64_t off = p * page;
    const int64_t len = (off + page <= nbytes) ? page : nbytes - off;
    memcpy((char*)dst + off, (const char*)src + off, (size_t)len);
  }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------
 * CSR SpMV, row split.  src/sparse/array/csr/spmv.cc:36-44
 *   for i in rows: sum = 0 (VAL_TY); for j_pos in pos[i]: sum += vals[j_pos]*x[crd[j_pos]]
 * Sequential left-to-right accumulation in the value type, alpha=1, beta=0.
 * ---------------------------------------------------------------------- */
#define SPMV_BODY(T)                                                        \
  do {                                                                      \
    const T* A = (const T*)vals; const T* X = (const T*)x; T* Y = (T*)y;    \
    for (int64_t i = 0; i < nrows; i++) {                                   \
      T sum = (T)0.0;                                                       \
      int64_t lo = ld_idx(indptr, pt, i), hi = ld_idx(indptr, pt, i + 1);   \
      for (int64_t p = lo; p < hi; p++) sum += A[p] * X[ld_idx(indices, it, p)]; \
      Y[i] = sum;                                                           \
    }                                                                       \
  } while (0)

int orc_spmv_csr(int vt, int it, int pt, int64_t nrows, const void* indptr,
                 const void* indices, const void* vals, const void* x, void* y)
{
  if (vt == 0) SPMV_BODY(float);
  else if (vt == 1) SPMV_BODY(double);
  else return ORC_EINVAL;
  return ORC_OK;
}

/* OpenMP variant.  src/sparse/array/csr/spmv_omp.cc:36-45
 * (#pragma omp parallel for schedule(monotonic:dynamic,128) over rows). */
#define SPMV_OMP_BODY(T)                                                    \
  do {                                                                      \
    const T* A = (const T*)vals; const T* X = (const T*)x; T* Y = (T*)y;    \
    _Pragma("omp parallel for schedule(monotonic:dynamic, 128)")            \
    for (int64_t i = 0; i < nrows; i++) {                                   \
      T sum = (T)0.0;                                                       \
      int64_t lo = ld_idx(indptr, pt, i), hi = ld_idx(indptr, pt, i + 1);   \
      for (int64_t p = lo; p < hi; p++) sum += A[p] * X[ld_idx(indices, it, p)]; \
      Y[i] = sum;                                                           \
    }                                                                       \
  } while (0)

int orc_spmv_csr_omp(int vt, int it, int pt, int64_t nrows, const void* indptr,
                     const void* indices, const void* vals, const void* x, void* y)
{
  if (vt == 0) SPMV_OMP_BODY(float);
  else if (vt == 1) SPMV_OMP_BODY(double);
  else return ORC_EINVAL;
  return ORC_OK;
}

/* ------------------------------------------------------------------------
 * CSR SpMM  Y = A @ X, X dense row-major (ncols x k, leading dimension ldx), Y (nrows x k, ldy).
 * src/sparse/array/csr/spmm.cc:37-50: zero the output block, then for each row i and each
 * nonzero kB of it (left to right), Y[i, j] += vals[kB] * X[crd[kB], j] for every j.
 * ---------------------------------------------------------------------- */
#define SPMM_BODY(T)                                                        \
  do {                                                                      \
    const T* A = (const T*)vals; const T* X = (const T*)x; T* Y = (T*)y;    \
    for (int64_t i = 0; i < nrows; i++)                                     \
      for (int64_t j = 0; j < k; j++) Y[i * ldy + j] = (T)0;                \
    for (int64_t i = 0; i < nrows; i++) {                                   \
      int64_t lo = ld_idx(indptr, pt, i), hi = ld_idx(indptr, pt, i + 1);   \
      for (int64_t p = lo; p < hi; p++) {                                   \
        int64_t c = ld_idx(indices, it, p);                                 \
        for (int64_t j = 0; j < k; j++) Y[i * ldy + j] += A[p] * X[c * ldx + j]; \
      }                                                                     \
    }                                                                       \
  } while (0)

int orc_spmm_csr(int vt, int it, int pt, int64_t nrows, int64_t k, const void* indptr,
                 const void* indices, const void* vals, const void* x, int64_t ldx, void* y, int64_t ldy)
{
  if (vt == 0) SPMM_BODY(float);
  else if (vt == 1) SPMM_BODY(double);
  else return ORC_EINVAL;
  return ORC_OK;
}

/* ------------------------------------------------------------------------
 * AXPBY (fused CG vector update).  src/sparse/linalg/axpby.cc:34-42
 *   val = a[0]/b[0]; if NEGATE val = -1*val;
 *   IS_ALPHA: y = val*x + y   else: y = x + val*y
 * a, b are 1-element arrays (the reference passes Legion futures).
 * ---------------------------------------------------------------------- */
#define AXPBY_BODY(T)                                                       \
  do {                                                                      \
    T* Y = (T*)y; const T* X = (const T*)x;                                 \
    T val = ((const T*)a)[0] / ((const T*)b)[0];                            \
    if (negate) val = (T)(-1) * val;                                        \
    for (int64_t i = 0; i < n; i++) {                                       \
      if (isalpha) Y[i] = val * X[i] + Y[i];                                \
      else         Y[i] = X[i] + val * Y[i];                                \
    }                                                                       \
  } while (0)

int orc_axpby(int vt, int64_t n, void* y, const void* x, const void* a, const void* b,
              int isalpha, int negate)
{
  if (vt == 0) AXPBY_BODY(float);
  else if (vt == 1) AXPBY_BODY(double);
  else return ORC_EINVAL;
  return ORC_OK;
}

/* ------------------------------------------------------------------------
 * dot / nrm2.  Call sites sparse/linalg.py:540,550 (r.dot(z), p.dot(q)) and :561
 * (np.linalg.norm(r)).  The arithmetic lives in cuNumeric branch-23.09, which is not
 * vendored under the reference; its published semantics are numpy's: sum_i x_i*y_i
 * and sqrt(sum_i x_i^2), result in the vector dtype.  Restated with an fp64
 * accumulator (the GPU path also accumulates in fp64 and rounds once at the end).
 * ---------------------------------------------------------------------- */
#define DOT_BODY(T)                                                         \
  do {                                                                      \
    const T* X = (const T*)x; const T* Y = (const T*)y; double s = 0.0;     \
    for (int64_t i = 0; i < n; i++) s += (double)X[i] * (double)Y[i];       \
    *((T*)out) = (T)s;                                                      \
  } while (0)

int orc_dot(int vt, int64_t n, const void* x, const void* y, void* out)
{
  if (vt == 0) DOT_BODY(float);
  else if (vt == 1) DOT_BODY(double);
  else return ORC_EINVAL;
  return ORC_OK;
}

int orc_nrm2(int vt, int64_t n, const void* x, void* out)
{
  double s = 0.0;
  if (vt == 0) { const float* X = (const float*)x; for (int64_t i = 0; i < n; i++) s += (double)X[i] * (double)X[i]; *((float*)out) = (float)sqrt(s); }
  else if (vt == 1) { const double* X = (const double*)x; for (int64_t i = 0; i < n; i++) s += X[i] * X[i]; *((double*)out) = sqrt(s); }
  else return ORC_EINVAL;
  return ORC_OK;
}

/* ------------------------------------------------------------------------
 * SpGEMM CSR x CSR -> CSR, phase 1: per-row nnz of the product.
 * src/sparse/array/csr/spgemm_csr_csr_csr.cc:61-81  (reference names: A = B*C;
 * here C = A*B).  Gustavson with `already_set` flags and an `index_list`; the count
 * is the number of distinct columns touched (cancellation zeros are kept because
 * the count is purely structural).
 * `row_nnz` is int64 (reference nnz_ty = uint64, util/typedefs.h:21).
 * ---------------------------------------------------------------------- */
int orc_spgemm_csr_nnz(int it, int pt, int64_t m, int64_t k, int64_t n,
                       const void* a_indptr, const void* a_indices,
                       const void* b_indptr, const void* b_indices, int64_t* row_nnz)
{
  (void)k;
  int64_t* index_list = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  unsigned char* already_set = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  if (!index_list || !already_set) { free(index_list); free(already_set); return ORC_ENOMEM; }
  for (int64_t i = 0; i < m; i++) {
    int64_t cnt = 0;
    for (int64_t kA = ld_idx(a_indptr, pt, i); kA < ld_idx(a_indptr, pt, i + 1); kA++) {
      int64_t kk = ld_idx(a_indices, it, kA);
      for (int64_t jB = ld_idx(b_indptr, pt, kk); jB < ld_idx(b_indptr, pt, kk + 1); jB++) {
        int64_t j = ld_idx(b_indices, it, jB);
        if (!already_set[j]) { index_list[cnt++] = j; already_set[j] = 1; }
      }
    }
    for (int64_t l = 0; l < cnt; l++) already_set[index_list[l]] = 0;
    row_nnz[i] = cnt;
  }
  free(index_list); free(already_set);
  return ORC_OK;
}

/* counts -> indptr.  sparse/base.py:30-48 (nnz_to_pos: cumsum of per-row counts;
 * pos[i] = {cumsum[i]-nnz[i], cumsum[i]-1}).  Plain exclusive scan here. */
int orc_nnz_to_indptr(int64_t m, const int64_t* row_nnz, int64_t* indptr)
{
  int64_t acc = 0;
  for (int64_t i = 0; i < m; i++) { indptr[i] = acc; acc += row_nnz[i]; }
  indptr[m] = acc;
  return ORC_OK;
}

/* ------------------------------------------------------------------------
 * SpGEMM phase 2: fill.  src/sparse/array/csr/spgemm_csr_csr_csr.cc:128-152
 * Per row: dense `workspace[j] += A_vals[kA]*B_vals[jB]` in (kA outer, jB inner)
 * order; output columns in FIRST-TOUCH order (index_list), not sorted; explicit
 * zeros from cancellation are kept.  c_indptr is int64; c_indices has width `it`.
 * ---------------------------------------------------------------------- */
#define SPGEMM_FILL_BODY(T)                                                 \
  do {                                                                      \
    const T* AV = (const T*)a_vals; const T* BV = (const T*)b_vals; T* CV = (T*)c_vals; \
    T* workspace = (T*)calloc((size_t)(n > 0 ? n : 1), sizeof(T));          \
    if (!workspace) { rc = ORC_ENOMEM; break; }                             \
    for (int64_t i = 0; i < m; i++) {                                       \
      int64_t cnt = 0;                                                      \
      for (int64_t kA = ld_idx(a_indptr, pt, i); kA < ld_idx(a_indptr, pt, i + 1); kA++) { \
        int64_t kk = ld_idx(a_indices, it, kA);                             \
        for (int64_t jB = ld_idx(b_indptr, pt, kk); jB < ld_idx(b_indptr, pt, kk + 1); jB++) { \
          int64_t j = ld_idx(b_indices, it, jB);                            \
          if (!already_set[j]) { index_list[cnt++] = j; already_set[j] = 1; } \
          workspace[j] += AV[kA] * BV[jB];                                  \
        }                                                                   \
      }                                                                     \
      int64_t pC = c_indptr[i];                                             \
      for (int64_t l = 0; l < cnt; l++) {                                   \
        int64_t j = index_list[l];                                          \
        already_set[j] = 0;                                                 \
        st_idx(c_indices, it, pC, j);                                       \
        CV[pC] = workspace[j];                                              \
        pC++;                                                               \
        workspace[j] = (T)0.0;                                              \
      }                                                                     \
    }                                                                       \
    free(workspace);                                                        \
  } while (0)

int orc_spgemm_csr_fill(int vt, int it, int pt, int64_t m, int64_t k, int64_t n,
                        const void* a_indptr, const void* a_indices, const void* a_vals,
                        const void* b_indptr, const void* b_indices, const void* b_vals,
                        const int64_t* c_indptr, void* c_indices, void* c_vals)
{
  (void)k;
  int rc = ORC_OK;
  int64_t* index_list = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  unsigned char* already_set = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  if (!index_list || !already_set) { free(index_list); free(already_set); return ORC_ENOMEM; }
  if (vt == 0) SPGEMM_FILL_BODY(float);
  else if (vt == 1) SPGEMM_FILL_BODY(double);
  else rc = ORC_EINVAL;
  free(index_list); free(already_set);
  return rc;
}

/* In-place per-row sort of (indices, vals) by column -- the canonical form both
 * sides are brought to before `indices` are compared bit-exactly (scipy's
 * csr_matmat and the reference both emit unsorted rows). Insertion
 * sort for short rows, qsort of an index permutation otherwise. */
typedef struct { int64_t c; int64_t p; } orc_pair;
static int orc_pair_cmp(const void* a, const void* b)
{
  int64_t ca = ((const orc_pair*)a)->c, cb = ((const orc_pair*)b)->c;
  return (ca > cb) - (ca < cb);
}

int orc_sort_rows(int vt, int it, int64_t m, const int64_t* indptr, void* indices, void* vals)
{
  size_t vs = vt == 0 ? 4 : 8;
  for (int64_t i = 0; i < m; i++) {
    int64_t lo = indptr[i], len = indptr[i + 1] - lo;
    if (len < 2) continue;
    orc_pair* pr = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)len);
    char* vtmp = (char*)malloc(vs * (size_t)len);
    if (!pr || !vtmp) { free(pr); free(vtmp); return ORC_ENOMEM; }
    for (int64_t l = 0; l < len; l++) { pr[l].c = ld_idx(indices, it, lo + l); pr[l].p = l; }
    qsort(pr, (size_t)len, sizeof(orc_pair), orc_pair_cmp);
    memcpy(vtmp, (char*)vals + vs * (size_t)lo, vs * (size_t)len);
    for (int64_t l = 0; l < len; l++) {
      st_idx(indices, it, lo + l, pr[l].c);
      memcpy((char*)vals + vs * (size_t)(lo + l), vtmp + vs * (size_t)pr[l].p, vs);
    }
    free(pr); free(vtmp);
  }
  return ORC_OK;
}

/* ------------------------------------------------------------------------
 * Row-block shard bounds.  sparse/csr.py:238-246 (tile = ceil(rows/num_procs)) and
 * sparse/partition.py:56-128 (CompressedImagePartition: nnz range of a contiguous
 * row block = [pos[lo].lo, pos[hi].hi]) / src/sparse/partition/fast_image_range.cc:27-35.
 * out[0..3] = row_lo, row_hi (exclusive), nnz_lo, nnz_hi (exclusive).
 * ---------------------------------------------------------------------- */
int orc_row_block(int pt, int64_t nrows, const void* indptr, int rank, int nranks, int64_t* out)
{
  if (nranks <= 0 || rank < 0 || rank >= nranks) return ORC_EINVAL;
  int64_t tile = (nrows + nranks - 1) / nranks;
  int64_t lo = (int64_t)rank * tile; if (lo > nrows) lo = nrows;
  int64_t hi = lo + tile; if (hi > nrows) hi = nrows;
  out[0] = lo; out[1] = hi;
  out[2] = ld_idx(indptr, pt, lo); out[3] = ld_idx(indptr, pt, hi);
  return ORC_OK;
}

/* Column window of a shard.  sparse/partition.py:139-208 (MinMaxImagePartition) /
 * src/sparse/partition/bounds_from_partitioned_coordinates.cc:29-40: [min col, max col]
 * over the shard's crd slice; empty slice -> lo=0, hi=-1. */
int orc_col_window(int it, const void* indices, int64_t nnz_lo, int64_t nnz_hi, int64_t* out)
{
  int64_t mn = INT64_MAX, mx = -1;
  for (int64_t p = nnz_lo; p < nnz_hi; p++) {
    int64_t c = ld_idx(indices, it, p);
    if (c < mn) mn = c;
    if (c > mx) mx = c;
  }
  if (mx < 0) { out[0] = 0; out[1] = -1; } else { out[0] = mn; out[1] = mx; }
  return ORC_OK;
}
