"""CPU oracle for the legate.sparse hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  ``legate.sparse_b200`` never does.

It wraps ``liboracle.so`` (``oracle.c``: C restatement of the reference leaf tasks) and
restates, in numpy, the two pieces of the path that are Python in the reference:

* ``cg``      -- sparse/linalg.py:499-565 (absolute ``tol``, convergence tested every
                ``conv_test_iters`` iterations, returns ``(x, iters)``)
* ``mmread``  -- src/sparse/io/mtx_to_coo.cc:47-137 + sparse/coo.py:233-347 (COO -> CSR by
                sorting on (row, col))

Parity status: PINNED against scipy.sparse 1.18.1 (the reference's own test oracle,
tests/integration/*.py) through tests/golden/*.npz, see tests/test_oracle_pinning.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _vt(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return 0
    if dtype == np.float64:
        return 1
    raise TypeError(f"oracle supports float32/float64 only, got {dtype}")


def _wide(arr: np.ndarray) -> int:
    if arr.dtype == np.int32:
        return 0
    if arr.dtype == np.int64:
        return 1
    raise TypeError(f"index arrays must be int32/int64, got {arr.dtype}")


def _p(arr: np.ndarray):
    return arr.ctypes.data_as(ctypes.c_void_p)


def _c(arr, dtype=None) -> np.ndarray:
    return np.ascontiguousarray(arr, dtype=dtype)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def first_touch_copy(arr: np.ndarray) -> np.ndarray:
    """A copy of `arr` whose pages are first touched by all OpenMP threads (bench.py CPU arms)."""
    arr = np.ascontiguousarray(arr)
    out = np.empty_like(arr)
    lib().orc_parallel_copy(_p(out), _p(arr), ctypes.c_int64(arr.nbytes))
    return out


def spmv(indptr, indices, data, x, omp: bool = False, out=None) -> np.ndarray:
    """y = A @ x; A and x are first promoted to a common dtype (sparse/csr.py:493).
    `out` (optional, right dtype/shape) avoids re-allocating y on every call (csr.py:509-513)."""
    indptr, indices = _c(indptr), _c(indices)
    dt = np.result_type(data.dtype, x.dtype)
    data, x = _c(data, dt), _c(x, dt)
    nrows = indptr.shape[0] - 1
    y = np.empty(nrows, dtype=dt) if out is None else out
    assert y.dtype == dt and y.shape == (nrows,) and y.flags.c_contiguous
    fn = lib().orc_spmv_csr_omp if omp else lib().orc_spmv_csr
    rc = fn(_vt(dt), _wide(indices), _wide(indptr), ctypes.c_int64(nrows), _p(indptr), _p(indices),
            _p(data), _p(x), _p(y))
    assert rc == 0, rc
    return y


def spmm(indptr, indices, data, X) -> np.ndarray:
    """Y = A @ X for a dense row-major X (ncols, k); spmm.cc:37-50, dtype promotion as csr.py:557."""
    indptr, indices = _c(indptr), _c(indices)
    dt = np.result_type(data.dtype, X.dtype)
    data, X = _c(data, dt), _c(X, dt)
    assert X.ndim == 2
    nrows, k = indptr.shape[0] - 1, X.shape[1]
    Y = np.empty((nrows, k), dtype=dt)
    rc = lib().orc_spmm_csr(_vt(dt), _wide(indices), _wide(indptr), ctypes.c_int64(nrows), ctypes.c_int64(k),
                            _p(indptr), _p(indices), _p(data), _p(X), ctypes.c_int64(k), _p(Y), ctypes.c_int64(k))
    assert rc == 0, rc
    return Y


def axpby(y, x, a, b, isalpha=True, negate=False) -> np.ndarray:
    """In-place fused update, sparse/linalg.py:479-496 / axpby.cc:34-42. a, b: 1-element arrays."""
    assert y.flags.c_contiguous and y.dtype == x.dtype
    a = _c(np.asarray(a).reshape(1), y.dtype)
    b = _c(np.asarray(b).reshape(1), y.dtype)
    x = _c(x)
    rc = lib().orc_axpby(_vt(y.dtype), ctypes.c_int64(y.shape[0]), _p(y), _p(x), _p(a), _p(b),
                         int(bool(isalpha)), int(bool(negate)))
    assert rc == 0, rc
    return y


def dot(x, y):
    x, y = _c(x), _c(y, x.dtype)
    out = np.empty(1, dtype=x.dtype)
    rc = lib().orc_dot(_vt(x.dtype), ctypes.c_int64(x.shape[0]), _p(x), _p(y), _p(out))
    assert rc == 0, rc
    return out


def nrm2(x):
    x = _c(x)
    out = np.empty(1, dtype=x.dtype)
    rc = lib().orc_nrm2(_vt(x.dtype), ctypes.c_int64(x.shape[0]), _p(x), _p(out))
    assert rc == 0, rc
    return out


def spgemm(a, b, shape_a, shape_b, sort_rows: bool = False):
    """C = A @ B for CSR triples (indptr, indices, data).  Returns (indptr int64, indices, data).

    Two passes exactly as the reference's CPU branch (sparse/csr.py:1390-1490): NNZ task ->
    nnz_to_pos (cumsum) -> fill task.  Rows come out in first-touch order unless ``sort_rows``.
    """
    a_ptr, a_idx, a_val = (_c(v) for v in a)
    b_ptr, b_idx, b_val = (_c(v) for v in b)
    dt = np.result_type(a_val.dtype, b_val.dtype)
    a_val, b_val = _c(a_val, dt), _c(b_val, dt)
    it = _wide(a_idx)
    assert _wide(b_idx) == it and _wide(a_ptr) == _wide(b_ptr)
    m, k = shape_a
    k2, n = shape_b
    assert k == k2
    L = lib()
    row_nnz = np.empty(m, dtype=np.int64)
    rc = L.orc_spgemm_csr_nnz(it, _wide(a_ptr), ctypes.c_int64(m), ctypes.c_int64(k), ctypes.c_int64(n),
                              _p(a_ptr), _p(a_idx), _p(b_ptr), _p(b_idx), _p(row_nnz))
    assert rc == 0, rc
    c_ptr = np.empty(m + 1, dtype=np.int64)
    L.orc_nnz_to_indptr(ctypes.c_int64(m), _p(row_nnz), _p(c_ptr))
    nnz = int(c_ptr[-1])
    c_idx = np.empty(nnz, dtype=a_idx.dtype)
    c_val = np.empty(nnz, dtype=dt)
    rc = L.orc_spgemm_csr_fill(_vt(dt), it, _wide(a_ptr), ctypes.c_int64(m), ctypes.c_int64(k),
                               ctypes.c_int64(n), _p(a_ptr), _p(a_idx), _p(a_val), _p(b_ptr), _p(b_idx),
                               _p(b_val), _p(c_ptr), _p(c_idx), _p(c_val))
    assert rc == 0, rc
    if sort_rows:
        rc = L.orc_sort_rows(_vt(dt), it, ctypes.c_int64(m), _p(c_ptr), _p(c_idx), _p(c_val))
        assert rc == 0, rc
    return c_ptr, c_idx, c_val


def row_block(indptr, rank: int, nranks: int):
    indptr = _c(indptr)
    out = np.empty(4, dtype=np.int64)
    rc = lib().orc_row_block(_wide(indptr), ctypes.c_int64(indptr.shape[0] - 1), _p(indptr), rank, nranks, _p(out))
    assert rc == 0, rc
    return tuple(int(v) for v in out)


def col_window(indices, nnz_lo: int, nnz_hi: int):
    indices = _c(indices)
    out = np.empty(2, dtype=np.int64)
    rc = lib().orc_col_window(_wide(indices), _p(indices), ctypes.c_int64(nnz_lo), ctypes.c_int64(nnz_hi), _p(out))
    assert rc == 0, rc
    return int(out[0]), int(out[1])


def cg(matvec, b, x0=None, tol=1e-8, maxiter=None, M=None, callback=None, conv_test_iters=25):
    """Restates sparse/linalg.py:499-565 op for op with the oracle's leaf functions.

    ``matvec(v) -> A@v``; ``M`` (optional) is ``z = M(r)``; default Identity = copy (:437-445).
    """
    b = np.asarray(b)
    n = b.shape[0]
    if maxiter is None:
        maxiter = n * 10
    x = np.zeros(n) if x0 is None else np.array(x0, copy=True)
    p = np.zeros(n)
    r = b - matvec(x)
    iters = 0
    rho = np.zeros(1, dtype=r.dtype)
    while iters < maxiter:
        z = r.copy() if M is None else M(r)
        rho1 = rho
        rho = dot(r, z)
        if iters == 0:
            p[:] = z
        else:
            axpby(p, z, rho, rho1, isalpha=False, negate=False)
        q = matvec(p)
        pq = dot(p, q)
        axpby(x, p, rho, pq, isalpha=True, negate=False)
        axpby(r, q, rho, pq, isalpha=True, negate=True)
        iters += 1
        if callback is not None:
            callback(x)
        if (iters % conv_test_iters == 0 or iters == (maxiter - 1)) and float(nrm2(r)[0]) < tol:
            break
    return x, iters


def mmread(path):
    """MatrixMarket coordinate reader, src/sparse/io/mtx_to_coo.cc:47-137.

    real / pattern / integer x general / symmetric; 1-based -> 0-based; for symmetric files the
    mirrored entry is emitted right after the original when row != col; values always float64,
    coordinates int64 (:28-29).  Returns (rows, cols, vals, (m, n)).
    """
    with open(path, "r") as f:
        header = f.readline().split()
        assert header[0] == "%%MatrixMarket" and header[1] == "matrix" and header[2] == "coordinate"
        field, symmetry = header[3], header[4]
        assert field in ("real", "pattern", "integer"), field
        assert symmetry in ("general", "symmetric"), symmetry
        symmetric = symmetry == "symmetric"
        line = f.readline()
        while line.lstrip().startswith("%"):
            line = f.readline()
        m, n, lines = (int(t) for t in line.split()[:3])
        rows, cols, vals = [], [], []
        for line in f:
            toks = line.split()
            if not toks:
                continue
            i, j = int(toks[0]), int(toks[1])
            if field == "pattern":
                v = 1.0
            elif field == "integer":
                v = float(int(toks[2]))
            else:
                v = float(toks[2])
            rows.append(i - 1); cols.append(j - 1); vals.append(v)
            if symmetric and i != j:
                rows.append(j - 1); cols.append(i - 1); vals.append(v)
    return (np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64),
            np.array(vals, dtype=np.float64), (m, n))


def coo_to_csr(rows, cols, vals, shape):
    """COO -> CSR by sorting on (row, col) (sparse/coo.py:233-347: SORT_BY_KEY then
    SORTED_COORDS_TO_COUNTS then nnz_to_pos).  Duplicates are assumed absent (coo.py:73-76)."""
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    counts = np.bincount(rows, minlength=shape[0]).astype(np.int64)
    indptr = np.zeros(shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, cols.astype(np.int64), vals
