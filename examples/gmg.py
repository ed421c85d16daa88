#!/usr/bin/env python
"""Geometric-multigrid-preconditioned CG on the 2-D Poisson problem -- the workload of the reference's
examples/gmg.py (V-cycle with weighted-Jacobi smoothing as the preconditioner M of linalg.cg).

Per level: restriction R (injection or full-weighting "linear"), prolongation P = R^T, Galerkin coarse
operator A_c = R A P (two SpGEMMs), smoother weight omega = (4/3) / rho(A D^-1) with rho estimated by power
iteration.  A V-cycle is: pre-smooth, residual, restrict (R.dot(..., spmv_domain_part=True) in the reference),
recurse, prolong, correct, post-smooth.

    python examples/gmg.py -n 512 -l 4 -m 200 [--package scipy]
"""
import argparse
import sys

import numpy as np

from common import select_package

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, default=128, help="grid is n x n")
ap.add_argument("-l", "--levels", type=int, default=3)
ap.add_argument("-m", "--maxiter", type=int, default=200)
ap.add_argument("-t", "--tol", type=float, default=1e-10)
ap.add_argument("-g", "--gridop", default="linear", choices=["linear", "injection"])
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--package", default="b200")
ap.add_argument("-v", "--verbose", action="store_true")
args = ap.parse_args()

name, timer, xp, sparse, linalg, on_device = select_package()
if on_device:
    import torch

    def rand(n, seed):
        return torch.from_numpy(np.random.default_rng(seed).random(n)).cuda()

    def norm(v):
        return float(torch.linalg.norm(v))

    def vdot(a, b):
        return float(torch.dot(a.reshape(-1), b.reshape(-1)))
else:

    def rand(n, seed):
        return np.random.default_rng(seed).random(n)

    def norm(v):
        return float(np.linalg.norm(v))

    def vdot(a, b):
        return float(np.dot(a.reshape(-1), b.reshape(-1)))


def poisson2d(n):
    """Standard 5-point Laplacian (4 on the diagonal, -1 to the four neighbours), row-major numbering."""
    N = n * n
    side = np.full(N - 1, -1.0)
    side[n - 1 :: n] = 0.0
    far = np.full(N - n, -1.0)
    M = sparse.diags([far, side, np.full(N, 4.0), side, far], [-n, -1, 0, 1, n], dtype=np.float64)
    return M.tocsc().T if name == "b200" else M.tocsr()


def restriction(fine_n, kind):
    """(coarse_n^2 x fine_n^2) restriction onto the even points of the fine grid."""
    cn = fine_n // 2
    ci, cj = np.divmod(np.arange(cn * cn, dtype=np.int64), cn)
    if kind == "injection":
        cols = (2 * ci) * fine_n + 2 * cj
        indptr = np.arange(cn * cn + 1, dtype=np.int64)
        return sparse.csr_matrix((np.ones(cn * cn), cols, indptr), shape=(cn * cn, fine_n * fine_n)), cn
    # full weighting: 3x3 stencil [1 2 1; 2 4 2; 1 2 1] / 16 centred on (2i, 2j), clipped at the boundary
    di, dj = np.meshgrid([-1, 0, 1], [-1, 0, 1], indexing="ij")
    w = (np.array([1.0, 2.0, 1.0])[:, None] * np.array([1.0, 2.0, 1.0])[None, :] / 16.0).reshape(-1)
    fi = (2 * ci)[:, None] + di.reshape(-1)[None, :]
    fj = (2 * cj)[:, None] + dj.reshape(-1)[None, :]
    ok = (fi >= 0) & (fi < fine_n) & (fj >= 0) & (fj < fine_n)
    cols = (fi * fine_n + fj)[ok]
    vals = np.broadcast_to(w, fi.shape)[ok]
    indptr = np.zeros(cn * cn + 1, dtype=np.int64)
    np.cumsum(ok.sum(axis=1), out=indptr[1:])
    return sparse.csr_matrix((vals, cols, indptr), shape=(cn * cn, fine_n * fine_n)), cn


def spectral_radius(A, iters=15, seed=1):
    v = rand(A.shape[1], seed).reshape(-1, 1)
    for _ in range(iters):
        v = A @ v
        v = v / norm(v)
    return vdot(v, A @ v)


class WeightedJacobi:
    def __init__(self, omega=4.0 / 3.0):
        self.base = omega
        self.levels = []

    def setup(self, A):
        d = A.diagonal()
        Dinv = sparse.eye(A.shape[0], dtype=A.dtype, format="csr")
        Dinv.data = 1.0 / d
        self.levels.append((self.base / spectral_radius(A @ Dinv), 1.0 / d))

    def pre(self, A, r, lvl):
        w, dinv = self.levels[lvl]
        return w * r * dinv

    def post(self, A, r, x, lvl):
        w, dinv = self.levels[lvl]
        return x + w * (r - A @ x) * dinv


class GMG:
    def __init__(self, A, n, levels, gridop):
        self.A, self.nlevels = A, levels
        self.smoother = WeightedJacobi()
        self.smoother.setup(A)
        self.ops = []
        for _ in range(levels):
            R, n = restriction(n, gridop)
            P = R.T.tocsr()
            A = R @ A @ P
            self.smoother.setup(A)
            self.ops.append((R, A, P))

    def cycle(self, A, r, lvl=0):
        if lvl == self.nlevels - 1:
            return self.smoother.pre(A, r, lvl)
        R, Ac, P = self.ops[lvl]
        x = self.smoother.pre(A, r, lvl)
        fine_r = r - A.dot(x)
        coarse_r = R.dot(fine_r, spmv_domain_part=True) if name == "b200" else R.dot(fine_r)
        x = x + P @ self.cycle(Ac, coarse_r, lvl + 1)
        return self.smoother.post(A, r, x, lvl)

    def as_preconditioner(self):
        return linalg.LinearOperator(self.A.shape, dtype=np.float64, matvec=lambda r: self.cycle(self.A, r))


timer.start()
A = poisson2d(args.n)
b = rand(args.n * args.n, args.seed)
print(f"Data creation time: {timer.stop():.3f} ms")
timer.start()
mg = GMG(A, args.n, args.levels, args.gridop)
M = mg.as_preconditioner()
print(f"GMG init time: {timer.stop():.3f} ms")
_ = M.matvec(xp.zeros(A.shape[1]))  # warm-up
residuals = []
cb = (lambda xk: residuals.append(norm(b - A @ xk))) if args.verbose else None
timer.start()
if name == "b200":
    x, iters = linalg.cg(A, b, tol=args.tol, maxiter=args.maxiter, M=M, callback=cb)
else:
    count = [0]

    def cb2(xk):
        count[0] += 1
        if cb:
            cb(xk)

    x, info = linalg.cg(A, b, rtol=0.0, atol=args.tol, maxiter=args.maxiter, M=M, callback=cb2)
    iters = count[0]
ms = timer.stop()
res = norm(b - A @ x)
print(f"{'Converged' if res < 10 * args.tol * max(1.0, norm(b)) else 'Stopped'} after {iters} iterations, |b - Ax| = {res:.3e}")
print(f"Solve Time: {ms:.3f} ms")
print(f"Iterations / sec: {iters / (ms / 1e3):.3f}")
if args.verbose:
    for k, r in enumerate(residuals):
        print(f"  iter {k + 1}: residual {r:.3e}")
