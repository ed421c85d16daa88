"""Construction helpers on the path of the examples: `diags`, `eye`, `identity`, `is_sparse_matrix`.

The reference builds these with cuNumeric ops + DIA->CSC conversion (sparse/module.py:96-246,
sparse/dia.py:175-249: entries equal to zero are dropped by the `data != 0` mask, dia.py:236).
They are matrix assembly, not the hot path: assembled on the host (vectorised numpy, no scipy
dependency in the arithmetic) and uploaded as a `csr_array`.
"""
from __future__ import annotations

import numpy as np

from .coo import coo_array
from .csr import csr_array


def is_sparse_matrix(o) -> bool:
    return isinstance(o, (csr_array, coo_array))


def _dia_to_csr(diagonals, offsets, shape, dtype):
    m, n = shape
    rows_l, cols_l, vals_l = [], [], []
    for d, k in zip(diagonals, offsets):
        k = int(k)
        length = min(m + min(k, 0), n - max(k, 0))
        if length <= 0:
            continue
        d = np.asarray(d)
        if d.ndim == 0:
            d = np.full(length, d)
        if d.shape[0] < length:
            raise ValueError(f"Diagonal length (index {k}: {d.shape[0]} at offset {k}) does not agree with "
                             f"array size ({m}, {n}).")
        d = d[:length]
        i = np.arange(length, dtype=np.int64) + max(-k, 0)
        j = i + k
        keep = d != 0  # dia.tocsc drops explicit zeros (reference dia.py:236)
        rows_l.append(i[keep]); cols_l.append(j[keep]); vals_l.append(d[keep])
    if rows_l:
        rows = np.concatenate(rows_l); cols = np.concatenate(cols_l); vals = np.concatenate(vals_l).astype(dtype)
    else:
        rows = cols = np.zeros(0, dtype=np.int64); vals = np.zeros(0, dtype=dtype)
    return coo_array((vals, (rows, cols)), shape=shape).tocsr()


class _dia_result:
    """What `diags(...)` returns when no format is requested: supports `.tocsr()`, `.tocsc().T`,
    `.todense()` -- the conversions the examples use (pde.py:163 `diags(...).tocsc().T`)."""

    def __init__(self, csr):
        self._csr = csr
        self.shape = csr.shape
        self.dtype = csr.dtype

    def tocsr(self, copy=False):
        return self._csr

    def tocsc(self, copy=False):
        return _csc_view(self._csr)

    def todense(self):
        return self._csr.todense()

    toarray = todense

    @property
    def T(self):
        return _dia_result(self._csr.T)


class _csc_view:
    """CSC of M represented by the CSR of M^T; `.T` hands that CSR back (reference csc.py:317-324)."""

    def __init__(self, csr_of_m):
        self._m = csr_of_m
        self.shape = csr_of_m.shape
        self.dtype = csr_of_m.dtype

    @property
    def T(self):
        return self._m.T

    def tocsr(self, copy=False):
        return self._m

    def todense(self):
        return self._m.todense()


def diags(diagonals, offsets=0, shape=None, format=None, dtype=None):
    """scipy.sparse.diags signature (reference sparse/module.py:96-218)."""
    if np.isscalar(offsets):
        if len(diagonals) == 0 or np.isscalar(diagonals[0]):
            diagonals = [np.atleast_1d(diagonals)]
        offsets = [offsets]
    else:
        diagonals = [np.atleast_1d(d) for d in diagonals]
    offsets = [int(o) for o in np.atleast_1d(offsets)]
    if len(diagonals) != len(offsets):
        raise ValueError("Different number of diagonals and offsets.")
    if shape is None:
        m = len(diagonals[0]) + abs(offsets[0])
        shape = (m, m)
    if dtype is None:
        dtype = np.result_type(*[np.asarray(d).dtype for d in diagonals])
    m, n = shape
    full = []
    for d, k in zip(diagonals, offsets):
        length = min(m + min(k, 0), n - max(k, 0))
        if length < 0:
            raise ValueError(f"Offset {k} (index) out of bounds")
        d = np.asarray(d)
        if d.shape[0] == 1 and length != 1:
            d = np.full(length, d[0])  # scalar broadcast, as scipy
        full.append(d)
    csr = _dia_to_csr(full, offsets, (int(m), int(n)), np.dtype(dtype))
    if format == "csr":
        return csr
    if format in (None, "dia"):
        return _dia_result(csr)
    if format == "csc":
        return _csc_view(csr)
    raise NotImplementedError(f"diags(format={format!r})")


def eye(m, n=None, k=0, dtype=np.float64, format=None):
    """Reference sparse/module.py:221-246 (returns a CSR identity-like matrix)."""
    if n is None:
        n = m
    m, n = int(m), int(n)
    length = max(0, min(m + min(k, 0), n - max(k, 0)))
    csr = _dia_to_csr([np.ones(length, dtype=dtype)], [k], (m, n), np.dtype(dtype))
    return csr


def identity(n, dtype=np.float64, format=None):
    return eye(n, n, dtype=dtype, format=format)
