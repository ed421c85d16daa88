"""legate.sparse_b200 -- B200-native drop-in for the legate.sparse hot path.

Same scipy.sparse-style surface as the reference package (sparse/__init__.py:21-31) for the one
path this project accelerates: `csr_array`/`csr_matrix` construction, `A @ x` / `A.dot(x)` (CSR
SpMV), `A @ B` (CSR x CSR SpGEMM) and `linalg.cg` with its axpby/dot/norm inner loop -- all executed by
hand-written sm_100a kernels in libb200sparse.so (no cuSPARSE, no Legion, no CPU fallback).
"""
from . import _lib  # noqa: F401  (loads libb200sparse.so; raises if it is missing)
from .coo import coo_array  # noqa: F401
from .csr import csr_array, csr_matrix, spgemm_csr_csr_csr  # noqa: F401
from .module import diags, eye, identity, is_sparse_matrix  # noqa: F401
from .runtime import runtime  # noqa: F401
from . import io, linalg  # noqa: F401

coo_matrix = coo_array

__version__ = "0.1.0"
