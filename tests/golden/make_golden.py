"""Generates tests/golden/golden.npz -- the pinned outputs the oracle (and the CUDA path) are checked against.

The reference ships NO golden vectors: its tests recompute scipy.sparse at test time on the five
.mtx fixtures (tests/integration/test_csr_dot.py:25-45, test_csr_spgemm.py:24-32, test_io.py:23-28)
and on seeded SPD matrices (test_cg_solve.py:23-36, utils/sample.py:25-44).  The reference itself
cannot be imported here (needs legate.core / Legion / cuNumeric), so this script freezes exactly
those scipy computations (scipy 1.18.1, numpy 2.3.5) into a small fixture.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.io as sio
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import MTX_FILES, sample_spd  # noqa: E402

out = {}
for name in MTX_FILES:
    key = name.split(".")[0]
    A = sio.mmread(os.path.join(HERE, name)).tocsr().astype(np.float64)
    A.sort_indices()
    n = A.shape[1]
    x = np.random.default_rng(0).random(n)
    out[f"{key}_indptr"] = A.indptr.astype(np.int64)
    out[f"{key}_indices"] = A.indices.astype(np.int64)
    out[f"{key}_data"] = A.data
    out[f"{key}_x"] = x
    out[f"{key}_y"] = A @ x
    out[f"{key}_y32"] = (A.astype(np.float32) @ x.astype(np.float32))
    # SpMM with a dense row-major operand, k = 8 (tests/integration/test_csr_dot.py's 2-D branch; csr.py:552-579)
    X = np.random.default_rng(1).random((n, 8))
    out[f"{key}_spmm_x"] = X
    out[f"{key}_spmm_y"] = A @ X
    C = A @ A
    C.sort_indices()
    out[f"{key}_c_indptr"] = C.indptr.astype(np.int64)
    out[f"{key}_c_indices"] = C.indices.astype(np.int64)
    out[f"{key}_c_data"] = C.data

# test.mtx known answer: A @ [1..5] = [14, 20, 0, 3, 8]
out["test_kat_y"] = np.array([14.0, 20.0, 0.0, 3.0, 8.0])

# CG on the reference's seeded SPD matrix (N=100, seed 471014) with tol=1e-8
Ad, xs = sample_spd(100, 0.1, 471014)
A = sp.csr_array(Ad)
y = A @ xs
xsol, info = spla.cg(A, y, rtol=0.0, atol=1e-8)
assert info == 0
out["cg100_y"] = y
out["cg100_x"] = xsol

# 5-point Laplacian of examples/pde.py:124-163, nx = ny = 18 -> N = 256
nx = ny = 18
dx = dy = 1.0 / (nx - 1)
a, g = 1.0 / dx**2, 1.0 / dy**2
c = -2.0 * a - 2.0 * g
diag_a = a * np.ones((nx - 2) * (ny - 2) - 1)
diag_a[nx - 3 :: nx - 2] = 0.0
diag_g = g * np.ones((nx - 2) * (ny - 3))
diag_c = c * np.ones((nx - 2) * (ny - 2))
L = sp.diags([diag_g, diag_a, diag_c, diag_a, diag_g], [-(nx - 2), -1, 0, 1, nx - 2], dtype=np.float64).tocsr()
L.eliminate_zeros()
L.sort_indices()
out["lap18_indptr"] = L.indptr.astype(np.int64)
out["lap18_indices"] = L.indices.astype(np.int64)
out["lap18_data"] = L.data
bl = np.ones(L.shape[0])
xl, info = spla.cg(L, bl, rtol=0.0, atol=1e-10)
assert info == 0
out["lap18_x"] = xl

np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
print("wrote golden.npz with", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")
