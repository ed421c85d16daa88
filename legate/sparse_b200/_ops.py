"""Leaf-op launchers: torch device tensors in, C-ABI calls out.

This is the layer that corresponds to the reference's "task builders" (free functions
`spmv` sparse/csr.py:863-968, `cg_axpby` sparse/linalg.py:479-496, `spgemm_csr_csr_csr`
sparse/csr.py:1317-1490): they marshal stores into a task launch; here they marshal device
pointers into libb200sparse.so.  Everything runs on the caller's current CUDA stream.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .runtime import idx_code, ptr, runtime, vt_code

L = _lib.lib


def _stream():
    return runtime.stream_ptr()


def _chk_dev(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            runtime.require_cuda("legate.sparse_b200 kernel launch")
            raise RuntimeError("operand is not a CUDA tensor; legate.sparse_b200 has no CPU fallback")


# ---- SpMV -----------------------------------------------------------------------------------------
class SpmvPlan:
    """Host handle + device buffer of a b2s SpMV plan (see include/b200sparse.h)."""

    def __init__(self, handle: int, buf: torch.Tensor):
        self.handle = handle
        self.buf = buf
        out = (_lib.c_i64 * 4)()
        _lib.check(L.b2s_spmv_plan_info(handle, out), "b2s_spmv_plan_info")
        self.config, self.rowgroup, self.uniform, self.scattered = int(out[0]), bool(out[1] & 1), bool(out[1] & 2), bool(out[1] & 4)
        self.short_rows, self.tma = bool(out[1] & 8), bool(out[1] & 16)
        self.tiles, self.lines_per_warp = int(out[2]), out[3] / 1000.0

    @property
    def chunks(self):
        """Row chunks for pipelined host<->device products: list of (tile_lo, tile_hi, row_lo, row_hi, col_lo, col_hi);
        empty when the plan is too small or not a TMA tile plan."""
        if getattr(self, "_chunks", None) is None:
            out = (_lib.c_i64 * (6 * 16))()
            n = ctypes.c_int(0)
            _lib.check(L.b2s_spmv_plan_chunks(self.handle, out, 16, ctypes.byref(n)), "b2s_spmv_plan_chunks")
            self._chunks = [tuple(int(out[6 * c + k]) for k in range(6)) for c in range(n.value)]
        return self._chunks

    def set_kernel(self, rowgroup: bool):
        _lib.check(L.b2s_spmv_plan_set_kernel(self.handle, int(bool(rowgroup))), "b2s_spmv_plan_set_kernel")
        self.rowgroup = bool(rowgroup)

    def set_flavor(self, flavor: int):
        """tools / tests: 0 generic, 1 uniform rows, 2 short rows (one lane per row)."""
        _lib.check(L.b2s_spmv_plan_set_flavor(self.handle, int(flavor)), "b2s_spmv_plan_set_flavor")
        self.uniform, self.short_rows = flavor == 1, flavor == 2

    @property
    def kernel_name(self) -> str:
        """The kernel this plan launches, for the bench line (derived from the plan, not a literal)."""
        if self.rowgroup:
            return "b2s::spmv_rowgroup_kernel"
        fl = "short-rows(one lane per row)" if self.short_rows else ("uniform-rows" if self.uniform else "generic")
        return f"b2s::spmv_{'tma' if self.tma else 'tile'}_kernel cfg {self.config} [{fl}]"

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                L.b2s_spmv_plan_destroy(h)
            except Exception:
                pass


def _plan_handle(plan):
    return None if plan is None else plan.handle


def spmv_plan(indptr: torch.Tensor, indices: torch.Tensor, shape, nnz: int, vdtype, tma_only: bool = False) -> SpmvPlan:
    """Build the SpMV plan for a CSR structure (tile boundaries + kernel choice).  `tma_only`: the plan must use the
    TMA tile kernel (needed by spmv_add / spmv_fused)."""
    _chk_dev(indptr, indices)
    vt = vt_code(vdtype)
    nrows, ncols = shape
    nbytes = int(L.b2s_spmv_plan_bytes(vt, nrows, nnz))
    buf = torch.empty(max(nbytes, 16) // 4, dtype=torch.int32, device=indptr.device)
    out = _lib.c_vp()
    _lib.check(L.b2s_spmv_plan_create_ex(vt, idx_code(indices.dtype), idx_code(indptr.dtype), nrows, ncols, nnz,
                                         ptr(indptr), ptr(indices), ptr(buf), _stream(), ctypes.byref(out),
                                         1 if tma_only else 0), "b2s_spmv_plan_create")
    return SpmvPlan(int(out.value), buf)


def spmv(indptr, indices, data, x, y, shape, plan=None):
    """y = A @ x (replaces task CSR_SPMV_ROW_SPLIT, sparse/csr.py:928-968)."""
    _chk_dev(indptr, indices, data, x, y)
    nrows, ncols = shape
    nnz = data.shape[0]
    assert x.dtype == data.dtype == y.dtype and x.is_contiguous() and y.is_contiguous()
    assert x.shape[0] == ncols and y.shape[0] == nrows
    _lib.check(L.b2s_spmv_csr(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows, ncols,
                              nnz, ptr(indptr), ptr(indices), ptr(data), ptr(x), ptr(y), _plan_handle(plan), _stream()),
               "b2s_spmv_csr")
    return y


def spmm(indptr, indices, data, X, Y, shape):
    """Y = A @ X for row-major dense X (ncols, k), Y (nrows, k) (replaces task SPMM_CSR_DENSE, sparse/csr.py:1151-1205).
    Rows of X / Y may be strided (leading dimension = stride(0)); the k entries of a row must be contiguous."""
    _chk_dev(indptr, indices, data, X, Y)
    nrows, ncols = shape
    assert X.dtype == data.dtype == Y.dtype and X.ndim == 2 and Y.ndim == 2
    k = X.shape[1]
    assert X.shape[0] == ncols and tuple(Y.shape) == (nrows, k)
    assert k <= 1 or (X.stride(1) == 1 and Y.stride(1) == 1)
    ldx = X.stride(0) if X.shape[0] > 1 else max(k, 1)
    ldy = Y.stride(0) if Y.shape[0] > 1 else max(k, 1)
    _lib.check(L.b2s_spmm_csr(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows, ncols,
                              data.shape[0], k, ptr(indptr), ptr(indices), ptr(data), ptr(X), ldx, ptr(Y), ldy,
                              _stream()), "b2s_spmm_csr")
    return Y


def spmv_tiles(indptr, indices, data, x, y, shape, plan, tile_lo: int, tile_hi: int):
    """Rows of tiles [tile_lo, tile_hi) of y = A @ x (x must be valid on that chunk's column window)."""
    _chk_dev(indptr, indices, data, x, y)
    nrows, ncols = shape
    _lib.check(L.b2s_spmv_csr_tiles(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows, ncols,
                                    data.shape[0], ptr(indptr), ptr(indices), ptr(data), ptr(x), ptr(y),
                                    _plan_handle(plan), tile_lo, tile_hi, _stream()), "b2s_spmv_csr_tiles")
    return y


def spmv_host(indptr, indices, data, x_host_ptr: int, y_host_ptr: int, x_dev, y_dev, shape, plan):
    """y_host = A @ x_host through the pipelined C entry point (returns when y_host is complete)."""
    _chk_dev(indptr, indices, data, x_dev, y_dev)
    nrows, ncols = shape
    _lib.check(L.b2s_spmv_csr_host(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows, ncols,
                                   data.shape[0], ptr(indptr), ptr(indices), ptr(data), x_host_ptr, y_host_ptr,
                                   ptr(x_dev), ptr(y_dev), _plan_handle(plan), _stream()), "b2s_spmv_csr_host")


def spmv_dot(indptr, indices, data, x, y, w, out, shape, plan):
    """y = A @ x and out[0] = w . y in one launch."""
    _chk_dev(indptr, indices, data, x, y, w, out)
    nrows, ncols = shape
    nnz = data.shape[0]
    assert x.dtype == data.dtype == y.dtype == w.dtype == out.dtype
    assert w.shape[0] == nrows and w.is_contiguous()
    ws = runtime.workspace()
    _lib.check(L.b2s_spmv_csr_dot(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows,
                                  ncols, nnz, ptr(indptr), ptr(indices), ptr(data), ptr(x), ptr(y), ptr(w),
                                  ptr(out), _plan_handle(plan), ptr(ws), _stream()), "b2s_spmv_csr_dot")
    return y


# ---- CG vector kernels --------------------------------------------------------------------------------
def axpby(y, x, a, b, isalpha=True, negate=False):
    _chk_dev(y, x, a, b)
    assert y.dtype == x.dtype == a.dtype == b.dtype and y.shape == x.shape
    assert y.is_contiguous() and x.is_contiguous()
    _lib.check(L.b2s_axpby(vt_code(y.dtype), y.numel(), ptr(y), ptr(x), ptr(a), ptr(b), int(bool(isalpha)),
                           int(bool(negate)), _stream()), "b2s_axpby")
    return y


def dot(x, y, out=None):
    _chk_dev(x, y)
    assert x.dtype == y.dtype and x.shape == y.shape and x.is_contiguous() and y.is_contiguous()
    if out is None:
        out = torch.empty(1, dtype=x.dtype, device=x.device)
    _lib.check(L.b2s_dot(vt_code(x.dtype), x.numel(), ptr(x), ptr(y), ptr(out), ptr(runtime.workspace()), _stream()),
               "b2s_dot")
    return out


def nrm2(x, out=None):
    _chk_dev(x)
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(1, dtype=x.dtype, device=x.device)
    _lib.check(L.b2s_nrm2(vt_code(x.dtype), x.numel(), ptr(x), ptr(out), ptr(runtime.workspace()), _stream()),
               "b2s_nrm2")
    return out


def cg_update_xr(x, r, p, q, rho, pq, rr_out):
    _chk_dev(x, r, p, q, rho, pq, rr_out)
    assert x.dtype == r.dtype == p.dtype == q.dtype == rho.dtype == pq.dtype == rr_out.dtype
    _lib.check(L.b2s_cg_update_xr(vt_code(x.dtype), x.numel(), ptr(x), ptr(r), ptr(p), ptr(q), ptr(rho), ptr(pq),
                                  ptr(rr_out), ptr(runtime.workspace()), _stream()), "b2s_cg_update_xr")
    return rr_out


def csr_diagonal(indptr, indices, data, nrows: int):
    _chk_dev(indptr, indices, data)
    out = torch.empty(nrows, dtype=data.dtype, device=data.device)
    _lib.check(L.b2s_csr_diagonal(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows,
                                  ptr(indptr), ptr(indices), ptr(data), ptr(out), _stream()), "b2s_csr_diagonal")
    return out


def copy(dst, src_ptr: int, n: int):
    """dst[:n] = *(src_ptr) -- src may be a peer-mapped (IPC) pointer."""
    _chk_dev(dst)
    _lib.check(L.b2s_copy(vt_code(dst.dtype), n, ptr(dst), src_ptr, _stream()), "b2s_copy")
    return dst


# ---- SpGEMM --------------------------------------------------------------------------------------------
def spgemm(a_indptr, a_indices, a_data, b_indptr, b_indices, b_data, shape_a, shape_b):
    """C = A @ B (replaces SPGEMM_CSR_CSR_CSR_GPU + scan_local_results_and_scale_pos,
    sparse/csr.py:1322-1389, 827-859). Returns (c_indptr int64, c_indices int32, c_data, info)."""
    _chk_dev(a_indptr, a_indices, a_data, b_indptr, b_indices, b_data)
    m, k = shape_a
    k2, n = shape_b
    assert k == k2
    assert a_indices.dtype == torch.int32 and b_indices.dtype == torch.int32, "SpGEMM needs int32 column indices"
    assert a_indptr.dtype == b_indptr.dtype and a_data.dtype == b_data.dtype
    dev = a_data.device
    pt = idx_code(a_indptr.dtype)
    vt = vt_code(a_data.dtype)
    st = _stream()
    scratch = torch.empty(int(L.b2s_spgemm_scratch_bytes(m, n)), dtype=torch.uint8, device=dev)
    c_indptr = torch.empty(m + 1, dtype=torch.int64, device=dev)
    info = (_lib.c_i64 * 3)()
    _lib.check(L.b2s_spgemm_csr_symbolic(pt, m, k, n, ptr(a_indptr), ptr(a_indices), ptr(b_indptr), ptr(b_indices),
                                         ptr(c_indptr), info, ptr(scratch), st), "b2s_spgemm_csr_symbolic")
    nnz, products, dense_rows = int(info[0]), int(info[1]), int(info[2])
    c_indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    c_data = torch.empty(nnz, dtype=a_data.dtype, device=dev)
    dense_bytes = int(L.b2s_spgemm_dense_bytes(vt, n, dense_rows))
    dense = torch.empty(dense_bytes, dtype=torch.uint8, device=dev) if dense_bytes else None
    _lib.check(L.b2s_spgemm_csr_numeric(vt, pt, m, k, n, ptr(a_indptr), ptr(a_indices), ptr(a_data), ptr(b_indptr),
                                        ptr(b_indices), ptr(b_data), ptr(c_indptr), ptr(c_indices), ptr(c_data),
                                        ptr(scratch), ptr(dense), dense_bytes, st), "b2s_spgemm_csr_numeric")
    return c_indptr, c_indices, c_data, {"nnz": nnz, "products": products, "dense_rows": dense_rows}


# ---- assembly -------------------------------------------------------------------------------------------------
def coo_to_csr(rows, cols, vals, nrows: int, ptr_dtype, idx_dtype):
    """COO triplets (device tensors) -> canonical CSR arrays (indptr, indices, vals) through b2s_coo_to_csr."""
    _chk_dev(rows, cols, vals)
    nnz = int(vals.shape[0])
    rows, cols = rows.to(idx_dtype).contiguous(), cols.to(idx_dtype).contiguous()
    vals = vals.contiguous()
    dev = vals.device
    indptr = torch.empty(nrows + 1, dtype=ptr_dtype, device=dev)
    indices = torch.empty(nnz, dtype=idx_dtype, device=dev)
    out_vals = torch.empty(nnz, dtype=vals.dtype, device=dev)
    pt = idx_code(ptr_dtype)
    scratch = torch.empty(int(L.b2s_convert_scratch_bytes(nrows, pt)), dtype=torch.uint8, device=dev)
    bad = _lib.c_i64(0)
    _lib.check(L.b2s_coo_to_csr(vt_code(vals.dtype), idx_code(idx_dtype), pt, nrows, nnz, ptr(rows), ptr(cols), ptr(vals),
                                ptr(indptr), ptr(indices), ptr(out_vals), ptr(scratch), ctypes.byref(bad), _stream()),
               "b2s_coo_to_csr")
    if bad.value:
        raise ValueError(f"{bad.value} triplets have a row index outside [0, {nrows})")
    return indptr, indices, out_vals


def csr_transpose(indptr, indices, data, shape):
    """CSR arrays of the transpose (same index widths) through b2s_csr_transpose."""
    _chk_dev(indptr, indices, data)
    nrows, ncols = shape
    nnz = int(data.shape[0])
    dev = data.device
    t_indptr = torch.empty(ncols + 1, dtype=indptr.dtype, device=dev)
    t_indices = torch.empty(nnz, dtype=indices.dtype, device=dev)
    t_vals = torch.empty(nnz, dtype=data.dtype, device=dev)
    pt = idx_code(indptr.dtype)
    scratch = torch.empty(int(L.b2s_convert_scratch_bytes(ncols, pt)), dtype=torch.uint8, device=dev)
    _lib.check(L.b2s_csr_transpose(vt_code(data.dtype), idx_code(indices.dtype), pt, nrows, ncols, nnz, ptr(indptr),
                                   ptr(indices.contiguous()), ptr(data.contiguous()), ptr(t_indptr), ptr(t_indices),
                                   ptr(t_vals), ptr(scratch), _stream()), "b2s_csr_transpose")
    return t_indptr, t_indices, t_vals


def spgemm_row_work(a_indptr, a_indices, b_indptr) -> torch.Tensor:
    """Products per row of A @ B (int64 device tensor of len m): the chunk planner's measure of work / output size."""
    _chk_dev(a_indptr, a_indices, b_indptr)
    assert a_indices.dtype == torch.int32 and a_indptr.dtype == b_indptr.dtype
    m = a_indptr.shape[0] - 1
    out = torch.empty(max(m, 1), dtype=torch.int64, device=a_indptr.device)
    _lib.check(L.b2s_spgemm_row_work(idx_code(a_indptr.dtype), m, ptr(a_indptr), ptr(a_indices), ptr(b_indptr), ptr(out),
                                     _stream()), "b2s_spgemm_row_work")
    return out[:m]


# ---- CUDA IPC / NVLink peer exchange -----------------------------------------------------------------------
def ipc_alloc(nbytes: int) -> int:
    out = _lib.c_vp()
    _lib.check(L.b2s_ipc_alloc(int(nbytes), ctypes.byref(out)), "b2s_ipc_alloc")
    return int(out.value)


def ipc_free(p: int) -> None:
    _lib.check(L.b2s_ipc_free(p), "b2s_ipc_free")


def ipc_export(p) -> bytes:
    if isinstance(p, torch.Tensor):
        _chk_dev(p)
        p = ptr(p)
    buf = ctypes.create_string_buffer(64)
    _lib.check(L.b2s_ipc_export(p, buf), "b2s_ipc_export")
    return buf.raw


def peer_allreduce(t: torch.Tensor, rank: int, peers) -> torch.Tensor:
    _chk_dev(t)
    n = len(peers)
    arr = (_lib.c_vp * n)(*peers)
    _lib.check(L.b2s_peer_allreduce(vt_code(t.dtype), rank, n, arr, ptr(t), t.numel(), _stream()), "b2s_peer_allreduce")
    return t


def peer_halo_exchange(x_local: torch.Tensor, rank: int, peers, sends, recv_peers) -> None:
    """sends: list of (peer, src_elem_off, dst_elem_off, count)."""
    _chk_dev(x_local)
    n = len(peers)
    arr = (_lib.c_vp * n)(*peers)
    flat = [int(v) for s in sends for v in s]
    desc = (_lib.c_i64 * max(len(flat), 1))(*flat)
    rp = (ctypes.c_int32 * max(len(recv_peers), 1))(*recv_peers)
    _lib.check(L.b2s_peer_halo_exchange(vt_code(x_local.dtype), rank, n, arr, ptr(x_local), len(sends), desc,
                                        len(recv_peers), rp, _stream()), "b2s_peer_halo_exchange")


def peer_push(x_local: torch.Tensor, rank: int, peers, sends, recv_peers, ctas_per_send: int = 8) -> None:
    """Fused-protocol push (no wait kernel): slices of x_local go to the peers' x buffers, `ctas_per_send` CTAs each;
    the arrival flags are consumed inside later `spmv_fused` launches."""
    _chk_dev(x_local)
    n = len(peers)
    arr = (_lib.c_vp * n)(*peers)
    flat = [int(v) for s in sends for v in s]
    desc = (_lib.c_i64 * max(len(flat), 1))(*flat)
    rp = (ctypes.c_int32 * max(len(recv_peers), 1))(*recv_peers)
    _lib.check(L.b2s_peer_push(vt_code(x_local.dtype), rank, n, arr, ptr(x_local), len(sends), desc,
                               len(recv_peers), rp, int(ctas_per_send), _stream()), "b2s_peer_push")


def peer_push_wait(rank: int, peers, recv_peers) -> None:
    """Wait on the current stream for the slices `recv_peers` pushed (b2s_peer_push) in the current exchange."""
    n = len(peers)
    arr = (_lib.c_vp * n)(*peers)
    rp = (ctypes.c_int32 * max(len(recv_peers), 1))(*recv_peers)
    _lib.check(L.b2s_peer_push_wait(rank, n, arr, len(recv_peers), rp, _stream()), "b2s_peer_push_wait")


def fuse_desc(ranges, n_free=None, flags=(), sends=(), acks=(), epoch_ctr=0, ticket=0, epoch_add=1, epoch_bump=1,
              expect=0, error=0, accumulate=False) -> "_lib.FuseDesc":
    """Build a b2s_fuse_desc.  ranges: [(tile_lo, tile_hi)]; flags: local arrival-word addresses;
    sends: [(src_ptr, dst_ptr, count, remote_flag_ptr, local_ack_ptr)]; acks: remote ack-word addresses."""
    d = _lib.FuseDesc()
    ranges = list(ranges)
    d.nranges = len(ranges)
    d.n_free = len(ranges) if n_free is None else int(n_free)
    for i, (lo, hi) in enumerate(ranges):
        d.ranges[2 * i], d.ranges[2 * i + 1] = int(lo), int(hi)
    d.n_flags, d.n_sends, d.n_acks = len(flags), len(sends), len(acks)
    d.accumulate = int(bool(accumulate))
    for i, f in enumerate(flags):
        d.flag[i] = int(f)
    for i, (src, dst, cnt, rflag, lack) in enumerate(sends):
        d.send_src[i], d.send_dst[i], d.send_count[i] = int(src), int(dst), int(cnt)
        d.send_flag[i], d.send_ack[i] = int(rflag), int(lack)
    for i, a in enumerate(acks):
        d.ack_out[i] = int(a)
    d.epoch_ctr = int(epoch_ctr) or None
    d.ticket = int(ticket) or None
    d.epoch_add, d.epoch_bump = int(epoch_add), int(epoch_bump)
    d.expect = int(expect)
    d.error = int(error) or None
    return d


def spmv_fused(indptr, indices, data, x, y, shape, plan, desc, w=None, dot_out=None):
    """y = A @ x (or y += with desc.accumulate) with the x exchange fused into the kernel; optional fused w . y."""
    _chk_dev(indptr, indices, data, x, y, w, dot_out)
    nrows, ncols = shape
    ws = ptr(runtime.workspace()) if w is not None else None
    _lib.check(L.b2s_spmv_csr_fused(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows, ncols,
                                    data.shape[0], ptr(indptr), ptr(indices), ptr(data), ptr(x), ptr(y),
                                    ptr(w) if w is not None else None, ptr(dot_out) if dot_out is not None else None,
                                    _plan_handle(plan), ws, ctypes.byref(desc), _stream()), "b2s_spmv_csr_fused")
    return y


def spmv_add(indptr, indices, data, x, y, shape, plan):
    """y += A @ x (TMA tile plans)."""
    _chk_dev(indptr, indices, data, x, y)
    nrows, ncols = shape
    _lib.check(L.b2s_spmv_csr_add(vt_code(data.dtype), idx_code(indices.dtype), idx_code(indptr.dtype), nrows, ncols,
                                  data.shape[0], ptr(indptr), ptr(indices), ptr(data), ptr(x), ptr(y),
                                  _plan_handle(plan), _stream()), "b2s_spmv_csr_add")
    return y


def peer_check(own_ptr: int) -> int:
    out = _lib.c_i64(0)
    _lib.check(L.b2s_peer_check(own_ptr, _stream(), ctypes.byref(out)), "b2s_peer_check")
    return int(out.value)


def ipc_open(handle: bytes) -> int:
    out = _lib.c_vp()
    _lib.check(L.b2s_ipc_open(handle, ctypes.byref(out)), "b2s_ipc_open")
    return int(out.value)


def ipc_close(p: int) -> None:
    _lib.check(L.b2s_ipc_close(p), "b2s_ipc_close")
