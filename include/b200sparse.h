/*
 * b200sparse.h -- C ABI of libb200sparse.so, the B200 (sm_100a) drop-in for the
 * legate.sparse hot path: CSR SpMV, CSR x CSR SpGEMM and the CG inner loop.
 *
 * What this boundary replaces.  The reference has no plain C ABI for the path: its .so
 * exports only `perform_registration()` and a projection-functor hook
 * (src/sparse/sparse_c.h:138-143) and every op is a Legate task
 * `static void X::gpu_variant(legate::TaskContext&)` looked up by opcode
 * (src/sparse/sparse_c.h:25-110 enum, e.g. CSR_SPMV_ROW_SPLIT, AXPBY,
 * SPGEMM_CSR_CSR_CSR_GPU) with stores fetched positionally
 * (src/sparse/array/csr/spmv_template.inl:77-84).  Each entry point below names the
 * task variant whose body it replaces; the Python side (legate/sparse_b200/*.py) plays
 * the role of the reference's task builders in sparse/csr.py and sparse/linalg.py.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in
 *    `_host`; the library never frees caller memory and never synchronises the stream
 *    unless the comment says "syncs";
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - return value: B2S_OK (0) or an error code; b2s_last_error() returns a
 *    thread-local message for the last failure (the reference aborts instead,
 *    src/sparse/util/cuda_help.h:51-74);
 *  - vt: value type 0=float32 1=float64; it: column-index width 0=int32 1=int64;
 *    pt: indptr width 0=int32 1=int64  (reference: util/dispatch.h:23-74);
 *  - CSR is scipy-style: indptr[nrows+1], indices[nnz], vals[nnz].  The reference's
 *    Rect<1> `pos` {lo,hi} (sparse/csr.py:186-203) is indptr[i], indptr[i+1]-1.
 */
#ifndef B200SPARSE_H
#define B200SPARSE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK            0
#define B2S_EINVAL        1   /* bad argument (null pointer, negative size, unknown type code) */
#define B2S_ECUDA         2   /* CUDA runtime error; message in b2s_last_error() */
#define B2S_EUNSUPPORTED  3   /* valid request this build does not implement */
#define B2S_ENOMEM        4   /* caller-provided buffer too small */

#define B2S_F32 0
#define B2S_F64 1
#define B2S_I32 0
#define B2S_I64 1

/* ---- library / device ------------------------------------------------------------- */
int         b2s_version(void);                 /* ABI version, currently 1 */
const char* b2s_last_error(void);              /* thread-local, never NULL */
/* out[0]=SM count, out[1]=L2 bytes, out[2]=cc major*10+minor, out[3]=max dyn smem/block.
 * Replaces Runtime.num_gpus / cudalibs handle bring-up (sparse/runtime.py:57-96,
 * src/sparse/cudalibs.cu:48-102). Host-side query; no stream. */
int         b2s_device_info(int device, int64_t* out4_host);

/* Size in bytes of the reduction workspace every *_dot / dot / nrm2 / cg_* call needs.
 * The caller allocates it once per (device, stream), zero-fills it once, and passes it
 * as `ws`; calls leave it zeroed again. */
int64_t     b2s_ws_bytes(void);

/* ---- CSR SpMV  (replaces CSRSpMVRowSplit::gpu_variant -> cusparseSpMV,
 *                 src/sparse/array/csr/spmv.cu:24-123,181-184) ---------------------- */

/* SpMV plan: the row-block / merge-path split of the (rows + nnz) work list into fixed-size
 * tiles, plus the kernel choice.  Stands in for the partitions the reference computes once
 * per store and caches (sparse/partition.py:56-128 CompressedImagePartition,
 * src/sparse/partition/fast_image_range.cu:27-54, bounds_from_partitioned_coordinates.cu).
 *   b2s_spmv_plan_bytes : size of the DEVICE buffer the caller provides (16-byte aligned);
 *   b2s_spmv_plan_create: fills it, samples the column locality of the matrix (mean number
 *       of distinct 128-byte x lines per 32 consecutive nonzeros) to choose the tile shape
 *       (streaming vs deep-gather) and kernel flavour, and returns a small HOST handle.
 *       syncs the stream once (the reference's partition tasks also block, fast_image_range.cu:27-54).
 *   b2s_spmv_plan_destroy: frees the host handle (never the device buffer).
 *   b2s_spmv_plan_info  : out[0]=tile config id, out[1]=bit0 row-group kernel, bit1 uniform-row path, bit2 scattered, bit3 short-row
 *                         (one lane per row) path, bit4 TMA tile kind, out[2]=tiles,
 *                         out[3]=1000*lines-per-warp statistic. */
int64_t     b2s_spmv_plan_tiles(int vt, int64_t nrows, int64_t nnz);
int64_t     b2s_spmv_plan_bytes(int vt, int64_t nrows, int64_t nnz);
int         b2s_spmv_plan_create(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                                 const void* indptr, const void* indices, void* plan_buf_dev,
                                 void* stream, void** plan_out);
/* flags: B2S_PLAN_TMA_ONLY = the caller will use b2s_spmv_csr_fused / b2s_spmv_csr_add, which need the TMA tile kernel */
#define B2S_PLAN_TMA_ONLY 1
int         b2s_spmv_plan_create_ex(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                                    const void* indptr, const void* indices, void* plan_buf_dev,
                                    void* stream, void** plan_out, int flags);
int         b2s_spmv_plan_destroy(void* plan);
int         b2s_spmv_plan_info(const void* plan, int64_t* out4_host);

/* y = A x  (alpha=1, beta=0 as spmv.cu:79-80).  `plan` may be NULL: then a plan-free
 * row-per-lane-group kernel is used (slower on short rows).  `plan` is the HOST handle
 * from b2s_spmv_plan_create (it must match vt/it/pt and the dimensions).  x has ncols
 * entries, y has nrows entries; y must not alias x.  The TMA-staged kernel needs
 * indptr/indices/vals 16-byte aligned (true for whole allocations); otherwise the
 * row-group kernel runs. */
int         b2s_spmv_csr(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                         const void* indptr, const void* indices, const void* vals,
                         const void* x, void* y, const void* plan, void* stream);

/* Row chunks for pipelined host<->device products.  A plan is cut into up to 16 chunks of tiles;
 * chunk c owns rows [row_lo,row_hi) and reads x only inside its column window [col_lo,col_hi)
 * (the reference's MinMaxImagePartition, sparse/partition.py:139-208, applied to row chunks of
 * one GPU).  b2s_spmv_plan_chunks: out = nchunks x {tile_lo,tile_hi,row_lo,row_hi,col_lo,col_hi}.
 * b2s_spmv_csr_tiles runs the tiles of one chunk, so x can still be arriving (H2D) for later
 * chunks while y of earlier chunks is already leaving (D2H). */
int         b2s_spmv_plan_chunks(const void* plan, int64_t* out_host, int max_chunks, int* nchunks_host);
int         b2s_spmv_csr_tiles(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                               const void* indptr, const void* indices, const void* vals,
                               const void* x, void* y, const void* plan,
                               int64_t tile_lo, int64_t tile_hi, void* stream);

/* SpMV with the x exchange FUSED into the kernel: one launch (a) pushes slices of the local x into neighbour GPUs'
 * x buffers over NVLink (remote stores by the first n_sends CTAs while all others already compute), (b) waits --
 * only before the first tile that reads remote columns -- for the slices the neighbours push here, (c)
 * acknowledges the previous exchange and (d) advances a device-side epoch, so the step replays from a CUDA graph
 * with no host-numbered arguments.  Replaces Legion's implicit halo copy + cusparseSpMV pair
 * (sparse/csr.py:928-968, partition.py:139-208).  All pointers in the descriptor are DEVICE addresses (local,
 * or peer-mapped through b2s_ipc_open where marked remote); the struct itself is read on the host.
 *   ranges      : tiles are visited range by range, ranges[2i], ranges[2i+1] = [tile_lo, tile_hi); the first
 *                 n_free ranges read only locally valid x
 *   flag[i]     : local arrival words (PeerHeader fuse_flag[src]); polled until >= epoch (bounded; *error = 1 on timeout)
 *   send_*[i]   : CTA i copies send_count[i] elements send_src[i] (local) -> send_dst[i] (remote) once
 *                 *send_ack[i] (local, written by that neighbour) >= epoch-1, then stores epoch to send_flag[i] (remote)
 *   ack_out[i]  : remote words that receive epoch-1 at kernel start (one per GPU that pushes into this one)
 *   epoch       : *epoch_ctr + epoch_add (device counter; the last CTA stores the epoch back if epoch_bump, using
 *                 *ticket for the election) or `expect` when epoch_ctr is NULL
 *   accumulate  : y += A x instead of y = A x (column-blocked shards; not with the fused dot)
 * w / dot_out / ws: as b2s_spmv_csr_dot, or all NULL. */
typedef struct b2s_fuse_desc {
  int32_t nranges, n_free;
  int64_t ranges[12];
  int32_t n_flags, n_sends, n_acks, accumulate;
  const void* flag[8];
  const void* send_src[4];
  void*       send_dst[4];
  int64_t     send_count[4];
  void*       send_flag[4];
  const void* send_ack[4];
  void*       ack_out[8];
  void*       epoch_ctr;
  void*       ticket;
  int32_t     epoch_add, epoch_bump;
  uint64_t    expect;
  void*       error;
} b2s_fuse_desc;
int         b2s_spmv_csr_fused(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                               const void* indptr, const void* indices, const void* vals,
                               const void* x, void* y, const void* w, void* dot_out, const void* plan,
                               void* ws, const b2s_fuse_desc* desc_host, void* stream);
/* y += A x (needs a TMA tile plan): one column block at a time -- the reduction step of the column-split SpMV (reference
 * sparse/csr.py:869-927, spmv_col_split_kernel src/sparse/array/csr/spmv.cu:125-153: x partitioned by columns, partial
 * products summed into y) and of the column-blocked shards of the multi-GPU all-gather exchange */
int         b2s_spmv_csr_add(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                             const void* indptr, const void* indices, const void* vals,
                             const void* x, void* y, const void* plan, void* stream);

/* y_host = A x_host for HOST vectors (the matrix stays resident): pipelined H2D / tiles / D2H over the plan's
 * chunks on internal copy streams (9 stages of 1-2-..-2-1 sixteenths, copy boundaries on 4 KB multiples);
 * x_dev / y_dev are caller-owned device scratch (ncols / nrows elements).  Pinned host memory gives true overlap.
 * Environment (read per call): B2S_PIPE_CHUNKS=n equal stages, B2S_PIPE_PATTERN="a,b,.." chunks per stage (sum 16),
 * B2S_PIPE_ALIGN=elements, B2S_PIPE_DIRECT=1 (tiles store y straight into mapped pinned memory), B2S_PIPE_TRACE=1.
 * syncs: y_host is complete on return. */
int         b2s_spmv_csr_host(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                              const void* indptr, const void* indices, const void* vals,
                              const void* x_host, void* y_host, void* x_dev, void* y_dev,
                              const void* plan, void* stream);

/* y = A x and *dot_out = sum_i w[i] * y[i] in one pass (CG: q = A p, pq = p.q;
 * sparse/linalg.py:549-550 fused).  w has nrows entries (for a row shard it is the
 * shard's slice of p).  dot_out: one value of type vt on the device. */
int         b2s_spmv_csr_dot(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                             const void* indptr, const void* indices, const void* vals,
                             const void* x, void* y, const void* w, void* dot_out,
                             const void* plan, void* ws, void* stream);

/* ---- SpMM: Y[nrows,k] = A @ X[ncols,k], dense operands row-major --------------------------
 * Replaces SpMMCSR::gpu_variant (src/sparse/array/csr/spmm.cu:25-110, cusparseSpMM ALG2 with
 * CUSPARSE_ORDER_ROW, alpha=1 beta=0) and the builder sparse/csr.py:1151-1205; CPU body spmm.cc:37-50.
 * ldx / ldy are the row strides of X / Y in elements (>= k).  Y is overwritten. */
int         b2s_spmm_csr(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz, int64_t k,
                         const void* indptr, const void* indices, const void* vals,
                         const void* X, int64_t ldx, void* Y, int64_t ldy, void* stream);

/* tools / tests: 0 = automatic (persistent TMA X-window kernel when the matrix has column locality, else fp64 ->
 * staged tile kernel, fp32 -> row kernel), 1 = one-row-per-lane-group kernel, 2 = staged tile kernel, 3 = staged tile
 * kernel with a synchronously loaded X window, 4 = TMA X-window kernel whenever the operand shape is eligible */
int         b2s_spmm_set_kernel(int kernel);

/* ---- CG vector kernels ---------------------------------------------------------------
 * b2s_axpby replaces AXPBY::gpu_variant (src/sparse/linalg/axpby.cu:25-62):
 *   val = a[0]/b[0]; negate -> -val; isalpha ? y = val*x + y : y = x + val*y
 * a_dev/b_dev are 1-element device arrays (the reference passes futures, linalg.py:479-496). */
int         b2s_axpby(int vt, int64_t n, void* y, const void* x, const void* a_dev,
                      const void* b_dev, int isalpha, int negate, void* stream);
/* r.dot(z), p.dot(q) and np.linalg.norm(r) (call sites sparse/linalg.py:540,550,561; the
 * arithmetic is cuNumeric's in the reference).  fp64 accumulation, deterministic order,
 * result (type vt) written to out_dev. */
int         b2s_dot (int vt, int64_t n, const void* x, const void* y, void* out_dev, void* ws, void* stream);
int         b2s_nrm2(int vt, int64_t n, const void* x, void* out_dev, void* ws, void* stream);
/* Fused CG step (sparse/linalg.py:553-555 + the next iteration's r.dot(z) with M = I):
 *   alpha = rho[0]/pq[0];  x += alpha p;  r -= alpha q;  *rr_out = r.r */
int         b2s_cg_update_xr(int vt, int64_t n, void* x, void* r, const void* p, const void* q,
                             const void* rho_dev, const void* pq_dev, void* rr_out_dev,
                             void* ws, void* stream);

/* diag_out[i] = A[i,i] (0 if absent). Replaces CSR_DIAGONAL / compute_diag_kernel
 * (src/sparse/array/csr/get_diagonal.cu:25-39), used by the GMG example's Jacobi smoother. */
int         b2s_csr_diagonal(int vt, int it, int pt, int64_t nrows, const void* indptr,
                             const void* indices, const void* vals, void* diag_out, void* stream);

/* ---- CSR x CSR -> CSR SpGEMM  (replaces SpGEMMCSRxCSRxCSRGPU::gpu_variant ->
 *      cusparseSpGEMM_{workEstimation,compute,copy}, src/sparse/array/csr/
 *      spgemm_csr_csr_csr.cu:33-272, and the NNZ/fill pair of the CPU branch,
 *      spgemm_csr_csr_csr.cc:26-154).  C[m,n] = A[m,k] * B[k,n].  int32 column indices.
 * Pass 1 (symbolic): c_indptr[m+1] (int64) = exclusive scan of per-row structural nnz.
 *   info_host[0] = nnz(C), info_host[1] = number of A*B products ("flops/2"),
 *   info_host[2] = rows with more than 1024 entries, i.e. beyond the 2048-entry shared-memory table (they take
 *   the dense accumulator in pass 2; size it with b2s_spgemm_dense_bytes).  syncs the stream
 *   (the reference also blocks here: sparse/csr.py:1442 `int(nnz)`).
 * Pass 2 (numeric): fills c_indices (int32) / c_vals; rows come out SORTED by column
 *   (canonical form; reference/scipy rows are unsorted, compare after sort_indices()).
 *   Structural zeros from cancellation are kept, as in the reference.
 * `scratch`: device buffer of b2s_spgemm_scratch_bytes(m, n) bytes shared by both passes.
 * `dense_ws`: device buffer of b2s_spgemm_dense_bytes(vt, n, info_host[2]) bytes (may be
 *   NULL when that is 0) -- the parallel form of the reference's per-thread `workspace` /
 *   `already_set` arrays (spgemm_csr_csr_csr.cc:100-118, _omp.cc:92-167). */
/* work_out[i] (int64, device) = A*B products of row i = upper bound of nnz(C[i,:]): lets a caller cut A into row chunks
 * whose output fits its memory budget and run the two passes chunk by chunk (a row slice of a CSR matrix is again a
 * CSR matrix: indptr slice rebased, contiguous indices/vals slice) -- how R-MAT scale 22 fits one GPU. */
int         b2s_spgemm_row_work(int pt, int64_t m, const void* a_indptr, const int32_t* a_indices,
                                const void* b_indptr, int64_t* work_out, void* stream);
int64_t     b2s_spgemm_scratch_bytes(int64_t m, int64_t n);
int64_t     b2s_spgemm_dense_bytes(int vt, int64_t n, int64_t dense_rows);
int         b2s_spgemm_csr_symbolic(int pt, int64_t m, int64_t k, int64_t n,
                                    const void* a_indptr, const int32_t* a_indices,
                                    const void* b_indptr, const int32_t* b_indices,
                                    int64_t* c_indptr, int64_t* info_host /*[3]*/,
                                    void* scratch, void* stream);
int         b2s_spgemm_csr_numeric(int vt, int pt, int64_t m, int64_t k, int64_t n,
                                   const void* a_indptr, const int32_t* a_indices, const void* a_vals,
                                   const void* b_indptr, const int32_t* b_indices, const void* b_vals,
                                   const int64_t* c_indptr, int32_t* c_indices, void* c_vals,
                                   void* scratch, void* dense_ws, int64_t dense_ws_bytes, void* stream);

/* ---- multi-GPU: NVLink peer-memory exchange (one process per GPU, CUDA IPC) -------------
 * Replaces the implicit Legion/Realm halo copies driven by MinMaxImagePartition
 * (sparse/partition.py:139-208) and the future-map reductions behind r.dot(z)/p.dot(q)
 * (sparse/linalg.py:540,550).  Each rank owns one buffer from b2s_ipc_alloc: the first
 * b2s_peer_header_bytes() bytes are flags/mailboxes, the x vector lives after them.  Handles
 * are 64 opaque bytes (cudaIpcMemHandle_t) exchanged by the caller (e.g. over
 * torch.distributed); `peers_host` is a host array of nranks device pointers (own buffer at
 * [rank], IPC-mapped peers elsewhere).  All spins are bounded; b2s_peer_check reports a timeout. */
int64_t     b2s_peer_header_bytes(void);
int         b2s_ipc_alloc(int64_t bytes, void** dev_ptr);          /* cudaMalloc + zero fill; syncs */
int         b2s_ipc_free(void* dev_ptr);
int         b2s_ipc_export(const void* dev_ptr, void* handle64_host);
int         b2s_ipc_open(const void* handle64_host, void** dev_ptr_out);
int         b2s_ipc_close(void* dev_ptr);
/* one-shot all-reduce (sum) of count <= 4 scalars of type vt, in place, identical result on every rank */
int         b2s_peer_allreduce(int vt, int rank, int nranks, void* const* peers_host, void* inout_dev,
                               int count, void* stream);
/* push slices of the local x into the neighbours' x buffers and wait for the slices they push here.
 * send_desc_host: nsends x {peer, src_elem_off, dst_elem_off, count}; recv_peers_host: nrecvs source ranks */
int         b2s_peer_halo_exchange(int vt, int rank, int nranks, void* const* peers_host,
                                   const void* x_local_dev, int nsends, const int64_t* send_desc_host,
                                   int nrecvs, const int32_t* recv_peers_host, void* stream);
/* fused protocol, all-gather style: slice i of the local x goes to peer i's x buffer with `ctas_per_send` CTAs per
 * destination; no wait kernel -- the consumers are b2s_spmv_csr_fused launches (epoch_add = 0) that poll the
 * arrival flag of the ONE source whose column block they multiply.  Device-side epochs (graph-replayable). */
int         b2s_peer_push(int vt, int rank, int nranks, void* const* peers_host, const void* x_local_dev,
                          int nsends, const int64_t* send_desc_host, int nrecvs,
                          const int32_t* recv_peers_host, int ctas_per_send, void* stream);
/* wait on `stream` for the slices the listed source ranks pushed with b2s_peer_push (stream-ordered after this rank's
 * own push of the same exchange): for consumers that need ALL of x before their first tile */
int         b2s_peer_push_wait(int rank, int nranks, void* const* peers_host, int nrecvs,
                               const int32_t* recv_peers_host, void* stream);
/* byte offset inside the peer header (fused protocol): which = 0 arrival flag of source rank idx, 1 error word,
 * 2 acknowledgement word of destination rank idx, 3 epoch counter, 4 ticket word */
int64_t     b2s_peer_header_offset(int which, int idx);
int         b2s_peer_check(void* own_buf_dev, void* stream, int64_t* error_out_host);   /* syncs */
/* dst[i] = src[i] for i in [0,n) elements of type vt where src is a (possibly peer) device
 * pointer; 128-bit loads when both are 16-byte aligned. */
int         b2s_copy(int vt, int64_t n, void* dst, const void* src, void* stream);

/* ---- assembly: COO triplets -> CSR, CSR -> CSR of the transpose ------------------------------
 * Replaces the reference's sort-by-key assembly (sparse/coo.py:233-347 -> src/sparse/sort/sort.cu:124-379,
 * sorted_coords_to_counts.cu:32, nnz_to_pos base.py:30-48) and its CSR->CSC->transpose route (sparse/csr.py:404-424)
 * with a counting sort by row (L2 atomics), a scan and a per-row bitonic sort by column: no global sort.  Unique
 * (row, col) pairs assumed, as in the reference (coo.py:73-76).  `scratch`: b2s_convert_scratch_bytes(nbuckets, pt)
 * bytes with nbuckets = nrows of the OUTPUT matrix.  Both calls sync the stream once. */
int64_t     b2s_convert_scratch_bytes(int64_t nbuckets, int pt);
int         b2s_coo_to_csr(int vt, int it, int pt, int64_t nrows, int64_t nnz, const void* rows, const void* cols,
                           const void* vals, void* indptr, void* indices, void* vals_out, void* scratch,
                           int64_t* bad_host, void* stream);
int         b2s_csr_transpose(int vt, int it, int pt, int64_t nrows, int64_t ncols, int64_t nnz,
                              const void* indptr, const void* indices, const void* vals,
                              void* t_indptr, void* t_indices, void* t_vals, void* scratch, void* stream);

/* ---- multi-GPU: NCCL communicator owned by the library -----------------------------------------
 * For hosts without torch.distributed (and for multi-node runs): the library binds NCCL at run time (dlopen of
 * libnccl.so.2, the one already in the process if any) and owns the communicator.  The reference uses NCCL
 * natively the same way (src/sparse/sort/sort.cu:163-322; communicator pre-initialised at sparse/runtime.py:84-87).
 *   b2s_comm_unique_id    : rank 0 fills 128 bytes (ncclUniqueId); the host ships them to every rank
 *   b2s_comm_init         : collective; communicator on the calling thread's current device
 *   b2s_allgather_x       : x_full[q*n_local ..] = rank q's x_local (equal shards; in place allowed) -- the
 *                           exchange before a row-sharded SpMV (sparse/csr.py:930-968)
 *   b2s_allreduce_scalars : in-place fp64 sum of the CG scalars (sparse/linalg.py:540,550,561)
 * Both collectives are enqueued on `stream` and do not synchronise. */
int         b2s_comm_nccl_version(void);      /* e.g. 22703; 0 = libnccl could not be loaded */
int         b2s_comm_unique_id(void* id128_host_out);
int         b2s_comm_init(int rank, int nranks, const void* nccl_unique_id_128B, void** comm_out);
int         b2s_comm_destroy(void* comm);
int         b2s_allgather_x(void* comm, int vt, const void* x_local, int64_t n_local, void* x_full, void* stream);
int         b2s_allreduce_scalars(void* comm, void* scalars_dev, int count, void* stream);

/* ---- measurement probe -------------------------------------------------------------------------------------
 * ngathers independent reads x[hash(i) mod ncols] (16 in flight per thread, nothing else read or written): the gather
 * rate the memory system sustains for the access pattern of a uniformly random CSR SpMV (BASELINE config 4).  Timed
 * by bench.py beside the R32 product; not used by any product path. */
int         b2s_probe_gather(int vt, int64_t ncols, int64_t ngathers, const void* x_dev, void* out_dev, void* stream);
/* cudaLimitMaxL2FetchGranularity of the current device (bytes L2 fetches from HBM per miss: 32 / 64 / 128; a driver
 * hint).  set_bytes = 0 reads it back only.  Relevant for scattered reads of vectors larger than L2. */
int         b2s_device_l2_fetch_granularity(int set_bytes, int64_t* current);

#ifdef __cplusplus
}
#endif
#endif /* B200SPARSE_H */
