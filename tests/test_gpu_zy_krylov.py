"""GPU tests of the Krylov solvers next to cg (cgs, bicg, bicgstab, gmres, lsqr, eigsh) driving the CUDA SpMV /
dot / nrm2 kernels through csr_array.  Systems, seeds and acceptance bars are the reference's
(tests/integration/test_cgs_solve.py:23-32, test_bicg_solve.py:23-45, test_gmres_solve.py:26-45,
test_lsqr_solve.py:23-32, test_eigsh.py:24-38)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from conftest import sample, sample_spd

import legate.sparse_b200 as sparse
from legate.sparse_b200 import linalg

pytestmark = pytest.mark.gpu


def _system(N=100, D=100):
    S, x = sample(N, D, 0.1, 471014)
    A = sparse.csr_array(np.asarray(S.todense()))
    return S, A, x, A @ x


def _shifted():
    # CGS / BiCGSTAB are rounding-lottery on the reference's indefinite sample (scipy's cgs stalls at 6e-5 after
    # 1000 iterations, its bicgstab diverges; the reference skips its own BiCGSTAB test): same matrix shifted to be
    # diagonally dominant, still non-symmetric
    S, _, x, _ = _system()
    S2 = sp.csr_array(S + 12.0 * sp.eye(100))
    return S2, sparse.csr_array(S2), x, S2 @ x


def test_cgs_solve():
    S2, A, x, y = _shifted()
    x_pred = linalg.cgs(A, y, tol=1e-8)
    assert isinstance(x_pred, np.ndarray)
    assert np.allclose(A @ x_pred, y, rtol=1e-5, atol=1e-6)
    ref, info = spla.cgs(S2, y, rtol=0, atol=1e-8)
    assert info == 0 and np.allclose(x_pred, ref, atol=1e-6)


def test_bicg_solve():
    S, A, x, y = _system()
    x_pred = linalg.bicg(A, y, tol=1e-8)
    assert np.allclose(A @ x_pred, y)
    # device vectors in -> device vector out
    yd = torch.from_numpy(y).cuda()
    xd = linalg.bicg(A, yd, tol=1e-8)
    assert isinstance(xd, torch.Tensor) and xd.is_cuda
    assert np.allclose(S @ xd.cpu().numpy(), y)


def test_bicgstab_solve():
    S2, A, x, y = _shifted()
    x_pred = linalg.bicgstab(A, y, tol=1e-8)
    assert np.allclose(A.dot(x_pred), y)
    ref, info = spla.bicgstab(S2, y, rtol=0, atol=1e-8)
    assert info == 0 and np.allclose(x_pred, ref, atol=1e-6)


def test_gmres_solve():
    S, A, x, y = _system()
    x_sci = spla.gmres(S, y, atol=1e-5, rtol=1e-5, maxiter=300, restart=20)[0]
    x_b200, info = linalg.gmres(A, y, atol=1e-5, tol=1e-5, maxiter=300)
    assert np.allclose(x_sci, x_b200, atol=1e-1)
    Ad, xs = sample_spd(80, 0.1, 3)
    P = sparse.csr_array(Ad)
    yp = Ad @ xs
    got, info = linalg.gmres(P, yp, tol=1e-10, restart=30)
    assert info == 0 and np.linalg.norm(Ad @ got - yp) <= 1.01e-10 * np.linalg.norm(yp)
    got32, info = linalg.gmres(P.astype(np.float32), yp.astype(np.float32), tol=1e-4, restart=30)
    assert got32.dtype == np.float32 and info == 0 and np.allclose(got32, xs, atol=1e-3)


def test_lsqr_solve():
    S, A, x, y = _system(1000, 500)
    res = linalg.lsqr(A, y, atol=1e-10, btol=1e-10)
    assert np.allclose(A @ res[0], y)
    ref = spla.lsqr(S, y, atol=1e-10, btol=1e-10)
    assert res[1] == ref[1] and abs(res[2] - ref[2]) <= 2 and np.allclose(res[0], ref[0], atol=1e-6)


def test_eigsh():
    S, _, _, _ = _system()
    Sd = np.asarray(S.todense())
    Sym = 0.5 * (Sd + Sd.T)
    A = sparse.csr_array(Sym)
    np.random.seed(0)
    vals, vecs = linalg.eigsh(A)
    for i, lamb in enumerate(vals):
        assert np.allclose(A @ vecs[:, i], lamb * vecs[:, i], atol=1e-3)
    exact = np.linalg.eigvalsh(Sym)
    assert np.allclose(vals, np.sort(exact[np.argsort(np.abs(exact))[-6:]]), atol=1e-8)
