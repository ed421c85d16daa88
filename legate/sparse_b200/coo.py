"""Minimal `coo_array`: the container `mmread` returns and the (data,(row,col)) constructor path.

The reference converts COO -> CSR with a distributed sort-by-key over NCCL (sparse/coo.py:233-347,
src/sparse/sort/sort.cu); that is matrix assembly, outside the SpMV/CG/SpGEMM hot path, so here it is a
host-side stable sort on (row, col) with the same result (duplicates are assumed absent, coo.py:73-76).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse

from .runtime import numpy_dtype, to_host


class coo_array:
    ndim = 2
    format = "coo"

    def __init__(self, arg, shape=None, dtype=None):
        if isinstance(arg, tuple) and len(arg) == 2:
            data, (row, col) = arg
            self._vals = np.asarray(to_host(data))
            self._i = np.asarray(to_host(row)).astype(np.int64, copy=False)
            self._j = np.asarray(to_host(col)).astype(np.int64, copy=False)
            if shape is None:
                shape = (int(self._i.max()) + 1, int(self._j.max()) + 1)
        elif scipy.sparse.issparse(arg):
            c = arg.tocoo()
            self._vals, self._i, self._j = c.data, c.row.astype(np.int64), c.col.astype(np.int64)
            shape = c.shape
        else:
            raise NotImplementedError
        if dtype is not None:
            self._vals = self._vals.astype(dtype)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = numpy_dtype(self._vals.dtype)

    @property
    def nnz(self):
        return int(self._vals.shape[0])

    @property
    def row(self):
        return self._i

    @property
    def col(self):
        return self._j

    @property
    def data(self):
        return self._vals

    def astype(self, dtype, casting="unsafe", copy=True):
        return coo_array((self._vals.astype(dtype, casting=casting, copy=copy), (self._i, self._j)), shape=self.shape)

    def tocoo(self, copy=False):
        return self

    def tocsr(self, copy=False):
        from .csr import csr_array

        order = np.lexsort((self._j, self._i))  # sort by (row, col), stable
        rows, cols, vals = self._i[order], self._j[order], self._vals[order]
        counts = np.bincount(rows, minlength=self.shape[0]).astype(np.int64)
        indptr = np.zeros(self.shape[0] + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])  # nnz_to_pos, reference base.py:30-48
        return csr_array((vals, cols, indptr), shape=self.shape)

    def todense(self):
        out = np.zeros(self.shape, dtype=self.dtype)
        np.add.at(out, (self._i, self._j), self._vals)
        return out

    toarray = todense

    def transpose(self, copy=False):
        return coo_array((self._vals, (self._j, self._i)), shape=(self.shape[1], self.shape[0]))

    T = property(transpose)
