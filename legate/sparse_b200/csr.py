"""`csr_array` -- the scipy.sparse-compatible CSR matrix of the hot path.

Mirrors the public surface of the reference class (sparse/csr.py:98-262 constructor forms,
:304-322 astype/copy, :426-440 to_scipy_sparse_csr, :442-582 dot, :799 __matmul__), but the
three Legate stores `pos`/`crd`/`vals` become three device tensors in plain scipy layout:

    indptr  int32 (int64 when nnz >= 2^31)   -- reference: Rect<1> pos, 16 B/row
    indices int32 (int64 when ncols >= 2^31) -- reference: int64 crd
    data    float32 / float64

so a 5-point Laplacian row costs 5*12 + 4 bytes instead of 5*16 + 16.  Dense vectors are either
numpy arrays (host; copied to the device and back around the kernel) or torch CUDA tensors
(device resident; the role cuNumeric ndarrays play in the reference).  Compute always happens in
libb200sparse.so; there is no CPU path.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import scipy.sparse
import torch

from . import _ops
from .runtime import (
    SUPPORTED_VALUE_DTYPES,
    is_device_array,
    numpy_dtype,
    runtime,
    to_device,
    to_host,
    torch_dtype,
)

_INT32_MAX = 2**31 - 1


def _force_wide() -> bool:
    # testing knob: keep 64-bit indices/indptr (exercises the int64 kernel instantiations)
    return os.environ.get("B2S_INDEX_WIDTH", "") == "64"


def cast_to_common_type(*args):
    """numpy.result_type promotion of the operands (reference sparse/utils.py:134-140)."""
    dts = [numpy_dtype(a.dtype) for a in args]
    common = np.result_type(*dts)
    return tuple(a.astype(common, copy=False) if numpy_dtype(a.dtype) != common else a for a in args)


class csr_array:
    """CSR matrix resident on one B200 (or, with no GPU visible, on the host for format logic)."""

    ndim = 2
    format = "csr"
    __array_priority__ = 10.1  # like scipy.sparse: `ndarray @ A` defers to A.__rmatmul__

    def __init__(self, arg, shape=None, dtype=None, copy=False):
        from .module import is_sparse_matrix

        self._plan = None
        self._plan_key = None
        if isinstance(arg, torch.Tensor) and arg.ndim == 2:
            arg = to_host(arg)
        if isinstance(arg, np.ndarray):
            # dense -> CSR: nonzero test is `!= 0`, columns ascending
            # (src/sparse/array/conv/dense_to_csr.cc:32-38,53-62)
            assert arg.ndim == 2
            shape = arg.shape
            mask = arg != 0
            counts = mask.sum(axis=1).astype(np.int64)
            indptr = np.zeros(shape[0] + 1, dtype=np.int64)
            np.cumsum(counts, out=indptr[1:])
            rows, cols = np.nonzero(mask)
            self._set_arrays(indptr, cols, arg[rows, cols])
        elif isinstance(arg, (scipy.sparse.csr_array, scipy.sparse.csr_matrix)):
            shape = arg.shape
            self._set_arrays(arg.indptr, arg.indices, arg.data, copy=True)
        elif isinstance(arg, tuple):
            if copy:
                raise NotImplementedError
            if shape is None:
                raise AssertionError("Cannot infer shape in this case.")
            if len(arg) == 2:
                # (data, (row, col)) goes through COO (reference csr.py:174-185)
                from .coo import coo_array

                data, (row, col) = arg
                res = coo_array((data, (row, col)), shape=shape).tocsr()
                self._indptr, self._indices, self._data = res._indptr, res._indices, res._data
                shape = res.shape
            elif len(arg) == 3:
                data, indices, indptr = arg
                assert indptr.shape[0] == shape[0] + 1
                self._set_arrays(indptr, indices, data)
            else:
                raise AssertionError
        elif is_sparse_matrix(arg):
            csr = arg.tocsr()
            if copy:
                csr = csr.copy()
            self._indptr, self._indices, self._data = csr._indptr, csr._indices, csr._data
            shape = csr.shape
        elif scipy.sparse.issparse(arg):
            s = arg.tocsr()
            shape = s.shape
            self._set_arrays(s.indptr, s.indices, s.data, copy=True)
        else:
            raise NotImplementedError

        assert shape is not None
        self.shape = tuple(int(i) for i in shape)
        assert self._indptr.shape[0] == self.shape[0] + 1
        if dtype is not None and numpy_dtype(dtype) != numpy_dtype(self._data.dtype):
            self._data = self._data.to(torch_dtype(dtype))
        self.dtype = numpy_dtype(self._data.dtype)

    # -- storage ------------------------------------------------------------------------------------
    def _set_arrays(self, indptr, indices, data, copy=False):
        nnz = int(data.shape[0])
        wide = _force_wide()
        ptr_dt = np.int64 if (wide or nnz > _INT32_MAX) else np.int32
        self._indptr = to_device(indptr, dtype=ptr_dt, copy=copy)
        if isinstance(indices, torch.Tensor):
            idx_dt = np.int64 if wide else np.int32
            if not wide and indices.numel() and int(indices.max()) > _INT32_MAX:
                idx_dt = np.int64
        else:
            indices = np.asarray(indices)
            idx_dt = np.int64 if (wide or (indices.size and int(indices.max()) > _INT32_MAX)) else np.int32
        self._indices = to_device(indices, dtype=idx_dt, copy=copy)
        self._data = to_device(data, copy=copy)

    # reference exposes `.data` / `.indices` as device arrays (csr.py:264-287); `.indptr` is added
    # because the layout here is plain scipy CSR.
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        self._data = to_device(value)
        self.dtype = numpy_dtype(self._data.dtype)
        # derived matrices (promoted copy, real expansion of complex values, transpose) follow the values
        self.__dict__.pop("_promo_cache", None)
        self.__dict__.pop("_expansion", None)
        self.__dict__.pop("_transposed", None)

    @property
    def indices(self):
        return self._indices

    @indices.setter
    def indices(self, value):
        self._indices = to_device(value)
        self._plan = None

    @property
    def indptr(self):
        return self._indptr

    @property
    def nnz(self) -> int:
        return int(self._data.shape[0])

    @property
    def device(self):
        return self._data.device

    @classmethod
    def _from_parts(cls, indptr, indices, data, shape):
        obj = cls.__new__(cls)
        obj._plan = None
        obj._plan_key = None
        obj._indptr, obj._indices, obj._data = indptr, indices, data
        obj.shape = tuple(int(i) for i in shape)
        obj.dtype = numpy_dtype(data.dtype)
        return obj

    @classmethod
    def make_empty(cls, shape, dtype):
        dev = runtime.device
        return cls._from_parts(
            torch.zeros(shape[0] + 1, dtype=torch.int32, device=dev),
            torch.zeros(0, dtype=torch.int32, device=dev),
            torch.zeros(0, dtype=torch_dtype(dtype), device=dev),
            shape,
        )

    # -- conversions ----------------------------------------------------------------------------------
    def astype(self, dtype, casting="unsafe", copy=True):
        dtype = np.dtype(dtype)
        if not copy and dtype == self.dtype:
            return self
        data = self._data.to(torch_dtype(dtype), copy=copy)
        indptr = self._indptr.clone() if copy else self._indptr
        indices = self._indices.clone() if copy else self._indices
        return csr_array._from_parts(indptr, indices, data, self.shape)

    def copy(self):
        return csr_array._from_parts(self._indptr.clone(), self._indices.clone(), self._data.clone(), self.shape)

    def conj(self, copy=True):
        if self.dtype.kind == "c":
            return csr_array._from_parts(self._indptr, self._indices, self._data.conj().resolve_conj(), self.shape)
        return self.copy() if copy else self   # real: the identity

    def tocsr(self, copy=False):
        return self.copy() if copy else self

    def tocoo(self, copy=False):
        from .coo import coo_array

        counts = (self._indptr[1:] - self._indptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(
            torch.arange(self.shape[0], dtype=torch.int64, device=self.device), counts
        )
        return coo_array((self._data.clone() if copy else self._data, (rows, self._indices)), shape=self.shape)

    def to_scipy_sparse_csr(self):
        return scipy.sparse.csr_array(
            (to_host(self._data), to_host(self._indices), to_host(self._indptr)), shape=self.shape, dtype=self.dtype
        )

    def todense(self, order=None, out=None):
        if order is not None:
            raise NotImplementedError
        res = self.to_scipy_sparse_csr().toarray()
        if out is not None:
            out[...] = res
            return out
        return res

    toarray = todense

    def transpose(self, copy=False):
        """CSR of the transpose (construction path: the reference reaches it through CSC aliasing,
        csc.py:317-324).  Counting sort by column with plain tensor ops on the device."""
        nrows, ncols = self.shape
        dev = self.device
        if (self._data.is_cuda and self._data.dtype in (torch.float32, torch.float64)
                and (self._indices.dtype == torch.int64 or nrows <= _INT32_MAX)):
            # the library's own kernels (b2s_csr_transpose, csrc/convert.cu): counting sort by column + per-row sort
            t_ptr, t_idx, t_val = _ops.csr_transpose(self._indptr, self._indices, self._data, self.shape)
            return csr_array._from_parts(t_ptr, t_idx, t_val, (ncols, nrows))
        counts = (self._indptr[1:] - self._indptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(nrows, dtype=torch.int64, device=dev), counts)
        cols = self._indices.to(torch.int64)
        order = torch.argsort(cols * max(nrows, 1) + rows)
        t_counts = torch.bincount(cols, minlength=ncols)
        t_indptr = torch.zeros(ncols + 1, dtype=torch.int64, device=dev)
        torch.cumsum(t_counts, 0, out=t_indptr[1:])
        idx_dt = torch.int64 if (_force_wide() or nrows > _INT32_MAX) else torch.int32
        ptr_dt = torch.int64 if (_force_wide() or self.nnz > _INT32_MAX) else torch.int32
        return csr_array._from_parts(t_indptr.to(ptr_dt), rows[order].to(idx_dt), self._data[order], (ncols, nrows))

    T = property(transpose)

    def diagonal(self, k=0):
        """Main diagonal as a device vector (reference csr.py:629-649; only k = 0 there too)."""
        if k != 0:
            raise NotImplementedError
        rows, cols = self.shape
        if self._data.is_cuda:
            d = _ops.csr_diagonal(self._indptr, self._indices, self._data, rows)
            return d[: min(rows, cols)]
        return to_device(self.to_scipy_sparse_csr().diagonal())

    # -- the hot path -----------------------------------------------------------------------------------
    def _get_plan(self, tma_only: bool = False):
        """Tile plan for the SpMV kernel, built once per structure and cached (the reference caches
        its image partitions per store the same way, sparse/partition.py:96-120).  `tma_only`: the plan must
        launch the TMA tile kernel (accumulating / exchange-fused products); sticky once requested."""
        from . import _lib

        tma_only = bool(tma_only or getattr(self, "_plan_tma_only", False))
        self._plan_tma_only = tma_only
        key = (self._indptr.data_ptr(), self._indices.data_ptr(), int(_lib.lib.b2s_spmv_get_config()),
               self.dtype.itemsize, self.nnz, tma_only)
        if self._plan is None or self._plan_key != key:
            self._plan = _ops.spmv_plan(self._indptr, self._indices, self.shape, self.nnz, self.dtype, tma_only=tma_only)
            self._plan_key = key
        return self._plan

    # -- column-split SpMV (reference csr.py:869-927 / spmv.cu:125-153: x partitioned, y reduced) -------------------
    _COL_BLOCK_BYTES = 40 << 20      # one block's slice of x: what stays L2-resident next to the matrix stream

    def _col_split(self, nblocks=None):
        """[(column block, plan)]: block q holds the entries whose column lies in [q*T, (q+1)*T), global column ids
        kept, so `y = sum_q A_q x` with every block gathering from ONE slice of x.  Built once per structure / values
        and cached.  The reference's column-split SpMV partitions x the same way and reduces the partial y with ADD."""
        n = self.shape[1]
        if nblocks is None:
            # scattered columns: one block per 40 MB of x (each block's slice stays in L2); otherwise the split is only
            # the partitioning the caller asked for -- two blocks (GMG's restriction: 514 it/s unsplit, 490 with five
            # blocks of a one-entry-per-row operator)
            scattered = bool(getattr(self._get_plan(), "scattered", False))
            nblocks = max(2, min(16, -(-n * self.dtype.itemsize // self._COL_BLOCK_BYTES))) if scattered else 2
        nblocks = max(1, min(int(nblocks), max(n, 1)))
        key = (self._indptr.data_ptr(), self._indices.data_ptr(), self._data.data_ptr(), self._data._version, self.nnz,
               nblocks)
        hit = self.__dict__.get("_colsplit")
        if hit is not None and hit[0] == key:
            return hit[1]
        T = -(-n // nblocks)
        idx = self._indices
        nrows = self.shape[0]
        counts = (self._indptr[1:] - self._indptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(nrows, device=idx.device, dtype=torch.int64), counts)
        blocks = []
        for q in range(nblocks):
            mask = (idx >= q * T) & (idx < min((q + 1) * T, n))
            nnz_q = int(mask.sum())
            if nnz_q == 0:
                continue
            cnt = torch.bincount(rows[mask], minlength=nrows)
            ip = torch.zeros(nrows + 1, dtype=torch.int64, device=idx.device)
            torch.cumsum(cnt, 0, out=ip[1:])
            B = csr_array._from_parts(ip.to(self._indptr.dtype), idx[mask].contiguous(), self._data[mask].contiguous(),
                                      self.shape)
            blocks.append((B, B._get_plan(tma_only=True)))
        self.__dict__["_colsplit"] = (key, blocks)
        return blocks

    def _dot_col_split(self, xd, y, nblocks=None):
        """y = A @ xd as a sum over column blocks: the first block writes y, the others accumulate (b2s_spmv_csr_add)."""
        blocks = self._col_split(nblocks)
        if not blocks:
            y.zero_()
            return y
        first = True
        for B, pl in blocks:
            if first:
                _ops.spmv(B._indptr, B._indices, B._data, xd, y, B.shape, plan=pl)
                first = False
            else:
                _ops.spmv_add(B._indptr, B._indices, B._data, xd, y, B.shape, pl)
        return y

    def _wants_col_split(self, plan) -> bool:
        """Scattered columns and an x larger than L2 can hold next to the matrix stream: gathers from one slice of x at a
        time hit L2, the unsplit product misses it (R32 fp64, x = 80 MB: 1.98 ms unsplit; weak-scaled fp32 shard, x =
        320 MB: 5.5 ms unsplit vs 1.65 ms in 8 blocks).  Only when the blocks keep a few entries per row."""
        if os.environ.get("B2S_COL_SPLIT", "auto") == "0":
            return False
        xbytes = self.shape[1] * self.dtype.itemsize
        if not getattr(plan, "scattered", False) or xbytes <= (48 << 20):
            return False
        nblocks = -(-xbytes // self._COL_BLOCK_BYTES)
        if self.nnz < 4 * nblocks * max(self.shape[0], 1):
            return False
        if "_colsplit" in self.__dict__:
            return True
        # the blocks are a second copy of the matrix (+ one indptr per block, + transient masks while they are cut):
        # only when that fits comfortably in what the device has free
        need = self.nnz * (self.dtype.itemsize + 4 + 9) + nblocks * (self.shape[0] + 1) * 16
        try:
            free, _total = torch.cuda.mem_get_info(self.device)
        except Exception:  # pragma: no cover - no usable device query: keep the unsplit product
            return False
        return need < 0.6 * free

    def _promoted(self, common):
        """A cast to the resolved dtype of (A, x) -- reference cast_to_common_type, csr.py:493 -- cached
        so a mixed-dtype SpMV does not re-cast the matrix on every call."""
        if self.dtype == common:
            return self
        cache = self.__dict__.setdefault("_promo_cache", {})
        hit = cache.get(common)
        stamp = (self._data.data_ptr(), self._data._version)   # in-place edits of .data keep the pointer
        if hit is None or hit[0] != stamp:
            cache[common] = (stamp, self.astype(common, copy=False))
        return cache[common][1]

    def _dot_host_pipelined(self, x: np.ndarray, out, plan):
        """y = A @ x for HOST vectors, software-pipelined over the plan's row chunks.

        Chunk c only reads x inside its column window (plan.chunks; the reference's MinMaxImagePartition idea
        applied to row chunks of one GPU), so its tiles can run as soon as x[:col_hi_c] has arrived, while the
        rest of x is still crossing PCIe and while y of earlier chunks is already on its way back.  For a
        banded matrix the product then costs about max(H2D, D2H) instead of H2D + kernel + D2H.  Pinned host
        arrays give true overlap; pageable ones still work (the copies just serialise)."""
        dev = self.device
        tdt = torch_dtype(self.dtype)
        st = self.__dict__.setdefault("_pipe", {})
        if st.get("key") != (self.shape, self.dtype):
            st.clear()
            st.update(key=(self.shape, self.dtype), xd=torch.empty(self.shape[1], dtype=tdt, device=dev),
                      yd=torch.empty(self.shape[0], dtype=tdt, device=dev))
        if out is None:
            out = np.empty(self.shape[0], dtype=self.dtype)
        _ops.spmv_host(self._indptr, self._indices, self._data, x.ctypes.data, out.ctypes.data, st["xd"], st["yd"],
                       self.shape, plan)
        return out

    # -- complex operands ------------------------------------------------------------------------------
    def _real_expansion(self, rdt):
        """Real (2m x 2n) CSR matrix E with E @ interleave(re x, im x) = interleave(re (A x), im (A x)): every
        complex entry a at (r, c) becomes the block [[re a, -im a], [im a, re a]] at rows 2r, 2r+1 / columns
        2c, 2c+1.  Built once per matrix with tensor ops and cached, so a complex SpMV is ONE launch of the real
        SpMV kernel on interleaved (re, im) storage -- which is exactly how complex vectors sit in memory."""
        hit = self.__dict__.get("_expansion")
        key = (self._data.data_ptr(), self._data._version, self._indptr.data_ptr(), np.dtype(rdt))
        if hit is not None and hit[0] == key:
            return hit[1]
        m, n = self.shape
        dev = self.device
        nnz = self.nnz
        tdt = torch_dtype(rdt)
        ptr64 = self._indptr.to(torch.int64)
        lens = ptr64[1:] - ptr64[:-1]
        rows = torch.repeat_interleave(torch.arange(m, dtype=torch.int64, device=dev), lens)
        q = torch.arange(nnz, dtype=torch.int64, device=dev) - ptr64[rows]      # position inside its row
        top = 4 * ptr64[rows] + 2 * q                                           # slot of (2r, 2c)
        bot = top + 2 * lens[rows]                                              # slot of (2r+1, 2c)
        cols = self._indices.to(torch.int64)
        data = self._data.to(torch.complex128 if np.dtype(rdt) == np.float64 else torch.complex64)
        re, im = data.real.to(tdt), data.imag.to(tdt)
        e_idx = torch.empty(4 * nnz, dtype=torch.int64, device=dev)
        e_val = torch.empty(4 * nnz, dtype=tdt, device=dev)
        e_idx[top], e_idx[top + 1], e_idx[bot], e_idx[bot + 1] = 2 * cols, 2 * cols + 1, 2 * cols, 2 * cols + 1
        e_val[top], e_val[top + 1], e_val[bot], e_val[bot + 1] = re, -im, im, re
        e_ptr = torch.empty(2 * m + 1, dtype=torch.int64, device=dev)
        e_ptr[0:2 * m:2] = 4 * ptr64[:-1]
        e_ptr[1:2 * m:2] = 4 * ptr64[:-1] + 2 * lens
        e_ptr[2 * m] = 4 * nnz
        wide = _force_wide()
        idx_dt = torch.int64 if (wide or 2 * n > _INT32_MAX) else torch.int32
        ptr_dt = torch.int64 if (wide or 4 * nnz > _INT32_MAX) else torch.int32
        E = csr_array._from_parts(e_ptr.to(ptr_dt), e_idx.to(idx_dt), e_val, (2 * m, 2 * n))
        self._expansion = (key, E)
        return E

    def _dot_complex(self, other, out):
        """A @ x / A @ X when the resolved dtype is complex (the reference dispatches its SpMV / SpMM tasks over
        complex64/128 too, src/sparse/util/dispatch.h).  No complex kernels: a real A with a complex operand is the
        real SpMM on the (re, im)-interleaved view of the operand (twice the columns); a complex A goes through its
        real expansion (`_real_expansion`)."""
        runtime.require_cuda("csr_array.dot")
        assert self.shape[1] == other.shape[0]
        m, n = self.shape
        common = np.result_type(self.dtype, numpy_dtype(other.dtype))
        rdt = np.dtype(np.float32) if common == np.complex64 else np.dtype(np.float64)
        if out is not None and numpy_dtype(out.dtype) != common:
            raise ValueError(f"Output type {numpy_dtype(out.dtype)} is not consistent with resolved dtype {common}")
        on_device = is_device_array(other) and other.is_cuda
        X = to_device(other, dtype=common)
        vector = X.ndim == 1
        k = 1 if vector else X.shape[1]
        Xr = torch.view_as_real(X.reshape(n, k))                                  # (n, k, 2) real view
        if self.dtype.kind != "c":
            Ar = self._promoted(rdt)
            Yr = torch.empty((m, 2 * k), dtype=Xr.dtype, device=Xr.device)
            _ops.spmm(Ar._indptr, Ar._indices, Ar._data, Xr.reshape(n, 2 * k), Yr, Ar.shape)
            Y = torch.view_as_complex(Yr.reshape(m, k, 2))
        else:
            E = self._real_expansion(rdt)
            if k == 1:
                yr = torch.empty(2 * m, dtype=Xr.dtype, device=Xr.device)
                _ops.spmv(E._indptr, E._indices, E._data, Xr.reshape(2 * n), yr, E.shape, plan=E._get_plan())
                Y = torch.view_as_complex(yr.reshape(m, 1, 2))
            else:
                Xe = Xr.permute(0, 2, 1).reshape(2 * n, k).contiguous()           # rows: re x_0, im x_0, re x_1, ...
                Ye = torch.empty((2 * m, k), dtype=Xr.dtype, device=Xr.device)
                _ops.spmm(E._indptr, E._indices, E._data, Xe, Ye, E.shape)
                Y = torch.view_as_complex(Ye.reshape(m, 2, k).permute(0, 2, 1).contiguous())
        Y = Y.reshape(m) if vector else Y.reshape(m, k)
        if out is not None:
            assert tuple(out.shape) == tuple(Y.shape)
            if isinstance(out, torch.Tensor):
                out.copy_(Y)
            else:
                out[...] = to_host(Y)
            return out
        return Y if on_device else to_host(Y)

    def dot(self, other, out=None, spmv_domain_part=False):
        """`A.dot(x)` / `A @ B`; see reference sparse/csr.py:442-582.

        * x 1-D or (n,1), numpy or torch CUDA tensor -> SpMV; result has x's array kind.
        * other dense 2-D (n,k) -> SpMM, dense (m,k) result of the same array kind.
        * other a csr_array -> SpGEMM (CSR x CSR -> CSR).
        `spmv_domain_part=True` (column-split SpMV, csr.py:869-927, spmv.cu:125-153): x is partitioned into column
        blocks, every block's partial product is reduced into y by the accumulating kernel (`_dot_col_split`); the
        same path is taken by itself when the columns are scattered and x exceeds L2 (`_wants_col_split`).
        """
        from .module import is_sparse_matrix

        if isinstance(other, csr_array):
            if out is not None:
                raise ValueError("Cannot provide out for CSRxCSR matmul.")
            assert self.shape[1] == other.shape[0]
            return spgemm_csr_csr_csr(*cast_to_common_type(self, other))
        if is_sparse_matrix(other) or scipy.sparse.issparse(other):
            other = np.asarray(other.todense())
            other_originally_sparse = True
        else:
            other_originally_sparse = False
        if not isinstance(other, (np.ndarray, torch.Tensor)):
            other = np.asarray(other)
        if other.ndim in (1, 2) and np.result_type(self.dtype, numpy_dtype(other.dtype)).kind == "c":
            res = self._dot_complex(other, out)
            if other_originally_sparse:
                return csr_array(np.asarray(to_host(res)).reshape(self.shape[0], -1))
            return res
        if other.ndim == 1 or (other.ndim == 2 and other.shape[1] == 1):
            runtime.require_cuda("csr_array.dot")
            assert self.shape[1] == other.shape[0]
            on_device = is_device_array(other) and other.is_cuda
            other_originally_2d = other.ndim == 2
            x = other.reshape(-1) if other_originally_2d else other
            if isinstance(x, torch.Tensor) and not x.is_contiguous():
                warnings.warn(
                    "CSR SpMV creating an implicit copy due to transformed x vector.",
                    category=RuntimeWarning,
                    stacklevel=2,
                )
            xdt = numpy_dtype(x.dtype)
            common = np.result_type(self.dtype, xdt)
            if common not in SUPPORTED_VALUE_DTYPES:
                raise NotImplementedError(
                    f"SpMV for resolved dtype {common} is not implemented (float32/float64 kernels only)"
                )
            A = self._promoted(common)
            if out is not None:
                odt = numpy_dtype(out.dtype)
                if odt != common:
                    raise ValueError(f"Output type {odt} is not consistent with resolved dtype {common}")
                if other_originally_2d:
                    assert tuple(out.shape) == (self.shape[0], 1)
                    assert not spmv_domain_part
                else:
                    assert tuple(out.shape) == (self.shape[0],)
            plan = A._get_plan()
            if (not on_device and isinstance(x, np.ndarray) and xdt == common and x.flags.c_contiguous
                    and (out is None or (isinstance(out, np.ndarray) and out.flags.c_contiguous))
                    and plan.chunks and not spmv_domain_part and os.environ.get("B2S_PIPELINE", "1") != "0"):
                # host vectors: stream x in / y out chunk by chunk, overlapping both PCIe directions with the kernel
                res = A._dot_host_pipelined(x, None if out is None else out.reshape(-1), plan)
                result = out if out is not None else (res.reshape(-1, 1) if other_originally_2d else res)
                if other_originally_sparse:
                    return csr_array(np.asarray(result).reshape(self.shape[0], -1))
                return result
            xd = to_device(x, dtype=common)
            direct = isinstance(out, torch.Tensor) and out.is_cuda and out.is_contiguous()
            y = out.reshape(-1) if direct else torch.empty(self.shape[0], dtype=torch_dtype(common), device=A.device)
            if spmv_domain_part or A._wants_col_split(plan):
                A._dot_col_split(xd, y)      # x partitioned by columns, partial products reduced into y
            else:
                _ops.spmv(A._indptr, A._indices, A._data, xd, y, A.shape, plan=plan)
            if out is None:
                result = y if on_device else to_host(y)
                if other_originally_2d:
                    result = result.reshape(-1, 1)
            else:
                if not direct:
                    if isinstance(out, torch.Tensor):
                        out.reshape(-1).copy_(y)
                    else:
                        flat = out.reshape(-1)
                        if flat.flags.c_contiguous and np.shares_memory(flat, out):
                            torch.from_numpy(flat).copy_(y)  # D2H straight into the caller's (pinned) buffer
                        else:
                            flat[...] = to_host(y)
                result = out
            if other_originally_sparse:
                return csr_array(np.asarray(to_host(result)).reshape(self.shape[0], -1))
            return result
        if other.ndim == 2:
            # SpMM, reference csr.py:552-579: C = A @ B with a dense row-major B; `out` must have the resolved dtype.
            runtime.require_cuda("csr_array.dot")
            assert self.shape[1] == other.shape[0]
            on_device = is_device_array(other) and other.is_cuda
            common = np.result_type(self.dtype, numpy_dtype(other.dtype))
            if common not in SUPPORTED_VALUE_DTYPES:
                raise NotImplementedError(
                    f"SpMM for resolved dtype {common} is not implemented (float32/float64 kernels only)"
                )
            A = self._promoted(common)
            k = other.shape[1]
            if out is not None:
                odt = numpy_dtype(out.dtype)
                if odt != common:
                    raise ValueError(f"Output type {odt} is not consistent with resolved dtype {common}")
                assert tuple(out.shape) == (self.shape[0], k)
            B = to_device(other, dtype=common)
            if k > 1 and B.stride(1) != 1:
                B = B.contiguous()  # transposed / column-strided operand: the kernel needs contiguous rows (csr.py:572-577)
            direct = isinstance(out, torch.Tensor) and out.is_cuda and (k <= 1 or out.stride(1) == 1)
            C = out if direct else torch.empty((self.shape[0], k), dtype=torch_dtype(common), device=A.device)
            _ops.spmm(A._indptr, A._indices, A._data, B, C, A.shape)
            if out is None:
                result = C if on_device else to_host(C)
            else:
                if not direct:
                    if isinstance(out, torch.Tensor):
                        out.copy_(C)
                    else:
                        out[...] = to_host(C)
                result = out
            if other_originally_sparse:
                return csr_array(np.asarray(to_host(result)))
            return result
        raise NotImplementedError("csr_array.dot: operand must be a vector, a dense 2-D array or a csr_array")

    def matvec(self, other):
        return self @ other

    def __matmul__(self, other):
        return self.dot(other)

    def __rmatmul__(self, other):
        """dense (i, m) @ A (m, n) -> dense (i, n)  (reference csr.py:801-822, rspmm csr.py:1208-1262).
        Computed as (A^T @ other^T)^T with the row-major SpMM kernel; A^T is built once and cached."""
        if not isinstance(other, (np.ndarray, torch.Tensor)):
            other = np.asarray(other)
        if other.ndim != 2:
            raise NotImplementedError
        assert other.shape[1] == self.shape[0]
        key = (self._indptr.data_ptr(), self._indices.data_ptr(), self._data.data_ptr(), self._data._version, self.nnz)
        if getattr(self, "_transposed", None) is None or self._transposed[0] != key:
            self._transposed = (key, self.transpose())
        At = self._transposed[1]
        if isinstance(other, torch.Tensor):
            return At.dot(other.t().contiguous()).t().contiguous()
        return np.ascontiguousarray(At.dot(np.ascontiguousarray(other.T)).T)

    def __mul__(self, other):
        if np.isscalar(other):
            return csr_array._from_parts(self._indptr, self._indices, self._data * other, self.shape)
        raise NotImplementedError

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1

    def balance(self):
        """No-op on a single device; kept for API parity (reference base.py:198-282 re-tiles the rows by
        nnz; the multi-GPU row plan lives in dist.py)."""
        return None

    def __repr__(self):
        return (f"<{self.shape[0]}x{self.shape[1]} legate.sparse_b200 csr_array, {self.nnz} stored elements, "
                f"dtype {self.dtype}, device {self.device}>")

    def __str__(self):
        return repr(self)


csr_matrix = csr_array


def spgemm_csr_csr_csr(A: csr_array, B: csr_array) -> csr_array:
    """C = A @ B (reference builder sparse/csr.py:1317-1490). Rows of C are sorted by column."""
    runtime.require_cuda("spgemm_csr_csr_csr")
    if A.dtype not in SUPPORTED_VALUE_DTYPES:
        raise NotImplementedError(f"SpGEMM for dtype {A.dtype} is not implemented")
    a_idx, b_idx = A._indices, B._indices
    if a_idx.dtype != torch.int32 or b_idx.dtype != torch.int32:
        if max(A.shape[1], B.shape[1]) > _INT32_MAX:
            raise NotImplementedError("SpGEMM needs column counts < 2^31")
        a_idx, b_idx = a_idx.to(torch.int32), b_idx.to(torch.int32)
    a_ptr, b_ptr = A._indptr, B._indptr
    if a_ptr.dtype != b_ptr.dtype:
        a_ptr, b_ptr = a_ptr.to(torch.int64), b_ptr.to(torch.int64)
    c_ptr, c_idx, c_val, info = _ops.spgemm(a_ptr, a_idx, A._data, b_ptr, b_idx, B._data, A.shape, B.shape)
    if info["nnz"] <= _INT32_MAX and not _force_wide():
        c_ptr = c_ptr.to(torch.int32)
    C = csr_array._from_parts(c_ptr, c_idx, c_val, (A.shape[0], B.shape[1]))
    C.spgemm_info = info
    return C


def spgemm_chunked(A: csr_array, B: csr_array, max_products: int = 1 << 31, keep: bool = False, on_chunk=None):
    """C = A @ B computed in ROW CHUNKS of A so that the output of one chunk fits a memory budget -- how BASELINE
    config 5 (R-MAT scale 22 squared: nnz(C) is hundreds of GB) runs on one 180 GB GPU.

    A cheap pass counts the products of every row (`b2s_spgemm_row_work`, an upper bound of the row's nnz); rows are
    cut greedily into ranges of <= `max_products` products (a single heavier row is its own chunk); each range is a
    CSR matrix in its own right (indptr slice rebased, contiguous indices / vals slice) and goes through the two-pass
    SpGEMM of `spgemm_csr_csr_csr`.  A chunk of C is handed to `on_chunk(row_lo, row_hi, C_chunk)` and dropped, or
    kept and concatenated when `keep` (only if everything fits).  The reference sizes its output with a blocking
    `int(nnz)` and a single allocation (sparse/csr.py:1442); it cannot run this case on one GPU either.

    Returns (C or None, stats) with stats = {chunks, products, nnz, rows_per_chunk, checksum (sum of all values,
    fp64), max_chunk_nnz}."""
    runtime.require_cuda("spgemm_chunked")
    assert A.shape[1] == B.shape[0]
    m = A.shape[0]
    a_idx = A._indices if A._indices.dtype == torch.int32 else A._indices.to(torch.int32)
    a_ptr, b_ptr = A._indptr, B._indptr
    if a_ptr.dtype != b_ptr.dtype:
        a_ptr, b_ptr = a_ptr.to(torch.int64), b_ptr.to(torch.int64)
    work = _ops.spgemm_row_work(a_ptr, a_idx, b_ptr)
    cum = torch.cumsum(work, 0).cpu().numpy() if m else np.zeros(0, dtype=np.int64)
    cuts = [0]
    while cuts[-1] < m:
        base = int(cum[cuts[-1] - 1]) if cuts[-1] > 0 else 0
        nxt = int(np.searchsorted(cum, base + int(max_products), side="right"))
        cuts.append(min(max(nxt, cuts[-1] + 1), m))
    del work
    parts = []
    stats = {"chunks": 0, "products": int(cum[-1]) if m else 0, "nnz": 0, "checksum": 0.0, "max_chunk_nnz": 0,
             "rows_per_chunk": []}
    ip_host = A._indptr.cpu().numpy() if m else np.zeros(1, dtype=np.int64)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        klo, khi = int(ip_host[lo]), int(ip_host[hi])
        Ac = csr_array._from_parts((A._indptr[lo : hi + 1] - klo).contiguous(), A._indices[klo:khi], A._data[klo:khi],
                                   (hi - lo, A.shape[1]))
        Cc = spgemm_csr_csr_csr(Ac, B)
        stats["chunks"] += 1
        stats["nnz"] += Cc.nnz
        stats["max_chunk_nnz"] = max(stats["max_chunk_nnz"], Cc.nnz)
        stats["rows_per_chunk"].append(hi - lo)
        stats["checksum"] += float(Cc._data.sum(dtype=torch.float64)) if Cc.nnz else 0.0
        if on_chunk is not None:
            on_chunk(lo, hi, Cc)
        if keep:
            parts.append(Cc)
        del Cc, Ac
    C = None
    if keep:
        ptrs, off = [torch.zeros(1, dtype=torch.int64, device=A.device)], 0
        for P_ in parts:
            ptrs.append(P_._indptr[1:].to(torch.int64) + off)
            off += P_.nnz
        indptr = torch.cat(ptrs)
        if off <= _INT32_MAX and not _force_wide():
            indptr = indptr.to(torch.int32)
        C = csr_array._from_parts(indptr, torch.cat([P_._indices for P_ in parts]) if parts else A._indices[:0],
                                  torch.cat([P_._data for P_ in parts]) if parts else A._data[:0], (m, B.shape[1]))
    return C, stats
