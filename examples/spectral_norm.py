#!/usr/bin/env python
"""Spectral-norm estimate by power iteration -- the workload of the reference's examples/spectral_norm.py: repeated
(n,1)-operand SpMV on a CSR matrix built from a dense array, checked against the same iteration on the dense array.

    python examples/spectral_norm.py [-n 100] [--package scipy]
"""
import argparse

import numpy as np

from common import select_package

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, default=100)
ap.add_argument("--package", default="b200")
args = ap.parse_args()
name, _, _, sparse, _, _ = select_package()


def normest(M, tol=1e-4, max_it=10):
    """||M||_2 of a symmetric positive semi-definite M (dense array or sparse matrix) by the power method."""
    rng = np.random.default_rng(15210)
    x = rng.random((M.shape[1], 1))
    y = np.asarray(M.dot(x))
    pnorm = np.sqrt(np.sum(y ** 2))
    x = y / pnorm
    res, it = 1.0, 0
    while res > tol and it < max_it:
        y = np.asarray(M.dot(x))
        ynorm = np.sqrt(np.sum(y ** 2))
        res = abs(pnorm - ynorm)
        pnorm = ynorm
        x = y / ynorm
        it += 1
    v = np.asarray(M.dot(x))
    return float(np.sqrt(np.sum(v ** 2)))


M = np.random.default_rng(15210).random((args.n, args.n))
A = sparse.csr_array(M)
est_sparse, est_dense = normest(A), normest(M)
assert np.isclose(est_sparse, est_dense), (est_sparse, est_dense)
print(f"[{name}] spectral norm estimate: {est_sparse:.10f} (dense: {est_dense:.10f})")
