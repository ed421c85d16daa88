"""The Krylov solvers that sit next to `cg` in the reference: `cgs`, `bicg`, `bicgstab`, `gmres`, `lsqr`, `eigsh`
(reference sparse/linalg.py:570-617, :622-667, :672-793, :798-838, :937-1413, :1416-1569).

What they share with the reference is the call surface and the result conventions:

* `cgs` / `bicg` / `bicgstab` return `x` only, `tol` is an ABSOLUTE bound on ||b - A x||_2, `M` must be None and
  `callback` is not supported (linalg.py:580-585);
* `gmres` returns `(x, info)`, right-preconditioned, restart default 20, `atol = max(atol, tol*||b||)`;
* `lsqr` returns scipy's 10-tuple `(x, istop, itn, r1norm, r2norm, anorm, acond, arnorm, xnorm, var)`;
* `eigsh` returns `(w, x)` (ascending) of the k extremal eigenpairs by thick-restart Lanczos.

Everything on the vector side runs on the device: A.matvec / A.rmatvec are the CSR SpMV kernel (through the
C ABI), inner products and norms are `b2s_dot` / `b2s_nrm2`, and the linear combinations of vectors are tensor
expressions with 1-element device tensors as coefficients (the reference writes them as cuNumeric expressions,
linalg.py:598-616) -- so an iteration needs at most one host round trip, the convergence test.  Unlike the
reference loops these honour `maxiter` (default 10 n) instead of spinning forever when a system does not converge.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _ops
from .runtime import is_device_array, numpy_dtype, runtime, to_device, to_host, torch_dtype


def _work_dtype(A, *arrays):
    dts = [numpy_dtype(a.dtype) for a in arrays if a is not None]
    if getattr(A, "dtype", None) is not None:
        dts.append(np.dtype(A.dtype))
    dt = np.result_type(*dts) if dts else np.dtype(np.float64)
    return dt if dt in (np.dtype(np.float32), np.dtype(np.float64)) else np.dtype(np.float64)


class _Space:
    """Inner products of the vector space the operator acts on.  For a single-GPU operator these are the plain
    `b2s_dot` / `b2s_nrm2` launches; an operator that carries an `allreduce` (a row shard of a distributed matrix,
    dist.dist_csr_array.as_linear_operator) gets its partial sums all-reduced, so the same solver code runs with
    every vector sharded by rows."""

    def __init__(self, op):
        self.reduce = getattr(op, "allreduce", None)

    def dot(self, x, y):
        d = _ops.dot(x, y)
        return d if self.reduce is None else self.reduce(d)

    def nrm2(self, x):
        if self.reduce is None:
            return _ops.nrm2(x)
        return torch.sqrt(self.reduce(_ops.dot(x, x)))

    def gram(self, basis, u):
        """basis (j, n) @ u (n): the j projections of u, summed over the shards."""
        h = basis @ u
        return h if self.reduce is None else self.reduce(h)


def _setup(A, b, x0, what):
    """-> (operator, its _Space, b on device (1-D), x on device (fresh copy), return-on-device flag)."""
    from .linalg import make_linear_operator

    runtime.require_cuda(what)
    assert len(b.shape) == 1 or (len(b.shape) == 2 and b.shape[1] == 1)
    op = make_linear_operator(A)
    on_device = is_device_array(b) and b.is_cuda
    dt = _work_dtype(op, b, x0)
    bd = to_device(b, dtype=dt).reshape(-1)
    if x0 is None:
        x = torch.zeros(op.shape[1], dtype=torch_dtype(dt), device=bd.device)
    else:
        x = to_device(x0, dtype=dt, copy=True).reshape(-1)
    return op, _Space(op), bd, x, on_device


def _global_rows(op, b):
    """System size for the default iteration cap (10 n): the global row count when b is a row shard."""
    return int(getattr(op, "global_shape", (b.shape[0],))[0])


def _finish(x, on_device):
    return x if on_device else to_host(x)


def _residual(op, b, x):
    return b - op.matvec(x)


def _confirm(op, sp, b, x, tol):
    """The short recurrences carry r along and it drifts away from b - A x in finite precision (CGS squares the
    drift).  Before a solver reports convergence the residual is recomputed from x: -> (true r, true ||r|| < tol).
    When it is not there yet the caller restarts its recurrence from the true residual."""
    r = _residual(op, b, x)
    return r, float(sp.nrm2(r)[0]) < tol


def _project_out(sp, basis, u):
    """u minus its components along the orthonormal rows of `basis`, by classical Gram-Schmidt applied TWICE, and
    the summed coefficients.  One pass (what the reference's gmres / eigsh do, linalg.py:753-755, :1420) loses
    orthogonality by a factor ||A v|| / ||u|| per step -- on a matrix like 10 I + E the Lanczos basis is garbage
    after ~18 steps -- the second pass restores it to rounding level ("twice is enough")."""
    h = sp.gram(basis, u)
    u = u - h @ basis
    h2 = sp.gram(basis, u)
    return u - h2 @ basis, h + h2


def _square_system(A, b):
    rows = A.local.shape[0] if hasattr(A, "local") else A.shape[0]   # b of a row shard holds the shard's rows
    assert len(A.shape) == 2 and A.shape[0] == A.shape[1] and b.shape[0] == rows


def _plain_only(M, callback):
    assert M is None, "preconditioning is not supported by this solver (reference linalg.py:580)"
    if callback is not None:
        raise NotImplementedError


def cgs(A, b, x0=None, tol=1e-5, maxiter=None, M=None, callback=None, atol=None):
    """Conjugate Gradient Squared (Sonneveld); reference linalg.py:570-617.  Returns x."""
    _square_system(A, b)
    _plain_only(M, callback)
    op, sp, b, x, on_device = _setup(A, b, x0, "linalg.cgs")
    maxiter = 10 * _global_rows(op, b) if maxiter is None else maxiter
    r = _residual(op, b, x)
    if float(sp.nrm2(r)[0]) < tol:
        return _finish(x, on_device)
    shadow = r.clone()              # fixed shadow residual
    p = r.clone()
    u = r.clone()
    rho = sp.dot(r, shadow)
    for _ in range(maxiter):
        Ap = op.matvec(p)
        alpha = rho / sp.dot(Ap, shadow)
        q = torch.addcmul(u, alpha, Ap, value=-1)            # q = u - alpha A p
        uq = u + q
        x.addcmul_(alpha, uq)                                # x += alpha (u + q)
        r = torch.addcmul(r, alpha, op.matvec(uq), value=-1)
        if float(sp.nrm2(r)[0]) < tol:
            r, done = _confirm(op, sp, b, x, tol)
            if done:
                break
            shadow, p, u = r.clone(), r.clone(), r.clone()
            rho = sp.dot(r, shadow)
            continue
        rho_next = sp.dot(r, shadow)
        beta = rho_next / rho
        rho = rho_next
        u = torch.addcmul(r, beta, q)                        # u = r + beta q
        p = torch.addcmul(u, beta, torch.addcmul(q, beta, p))  # p = u + beta (q + beta p)
    return _finish(x, on_device)


def bicg(A, b, x0=None, tol=1e-5, maxiter=None, M=None, callback=None, atol=None):
    """BiConjugate Gradient; needs A.rmatvec (the cached transpose for a csr_array).  Reference
    linalg.py:622-667.  Returns x."""
    _square_system(A, b)
    _plain_only(M, callback)
    op, sp, b, x, on_device = _setup(A, b, x0, "linalg.bicg")
    maxiter = 10 * _global_rows(op, b) if maxiter is None else maxiter
    r = _residual(op, b, x)
    if float(sp.nrm2(r)[0]) < tol:
        return _finish(x, on_device)
    rs = b.clone()                  # shadow residual b - A^T 0 (linalg.py:646-647)
    p = r.clone()
    ps = rs.clone()
    rho = sp.dot(rs, r)
    for _ in range(maxiter):
        Ap = op.matvec(p)
        alpha = rho / sp.dot(ps, Ap)
        x.addcmul_(alpha, p)
        r = torch.addcmul(r, alpha, Ap, value=-1)
        rs = torch.addcmul(rs, alpha, op.rmatvec(ps), value=-1)
        if float(sp.nrm2(r)[0]) < tol:
            r, done = _confirm(op, sp, b, x, tol)
            if done:
                break
            rs, p, ps = r.clone(), r.clone(), r.clone()
            rho = sp.dot(rs, r)
            continue
        rho_next = sp.dot(rs, r)
        beta = rho_next / rho
        rho = rho_next
        p = torch.addcmul(r, beta, p)
        ps = torch.addcmul(rs, beta, ps)
    return _finish(x, on_device)


def bicgstab(A, b, x0=None, tol=1e-5, maxiter=None, M=None, callback=None, atol=None):
    """BiCGSTAB (van der Vorst) with the reference's restart safeguard: when r.rhat collapses below 1e-8 the
    shadow vector and the search direction are reset to the current residual (linalg.py:798-838).  Returns x."""
    _plain_only(M, callback)
    op, sp, b, x, on_device = _setup(A, b, x0, "linalg.bicgstab")
    maxiter = 10 * _global_rows(op, b) if maxiter is None else maxiter
    r = _residual(op, b, x)
    if float(sp.nrm2(r)[0]) < tol:
        return _finish(x, on_device)
    shadow = r.clone()
    p = r.clone()
    rho = sp.dot(r, shadow)
    for _ in range(maxiter):
        Ap = op.matvec(p)
        alpha = rho / sp.dot(Ap, shadow)
        s = torch.addcmul(r, alpha, Ap, value=-1)
        if float(sp.nrm2(s)[0]) < tol:
            x.addcmul_(alpha, p)
            r, done = _confirm(op, sp, b, x, tol)
            if done:
                break
            shadow, p = r.clone(), r.clone()
            rho = sp.dot(r, shadow)
            continue
        As = op.matvec(s)
        omega = sp.dot(As, s) / sp.dot(As, As)
        x.addcmul_(alpha, p).addcmul_(omega, s)
        r = torch.addcmul(s, omega, As, value=-1)
        rho_next = sp.dot(r, shadow)
        rnorm, rho_host = torch.cat([sp.nrm2(r), rho_next]).tolist()   # one host round trip
        if rnorm < tol:
            r, done = _confirm(op, sp, b, x, tol)
            if done:
                break
            rho_host = 0.0                      # not there yet: restart from the true residual
        if abs(rho_host) < 1e-8:
            shadow = r.clone()
            p = r.clone()
            rho = sp.dot(r, shadow)
            continue
        beta = (alpha / omega) * (rho_next / rho)
        rho = rho_next
        p = torch.addcmul(r, beta, torch.addcmul(p, omega, Ap, value=-1))   # p = r + beta (p - omega A p)
    return _finish(x, on_device)


def gmres(A, b, x0=None, tol=1e-5, restart=None, maxiter=None, M=None, callback=None, atol=None,
          callback_type=None, conv_test_iters=25):
    """Restarted GMRES with right preconditioning; reference linalg.py:672-793.  Returns (x, info): info = 0 on
    convergence, else the number of iterations spent.  The Krylov basis lives on the device as `restart` contiguous
    rows; the small least-squares problem H y = e is solved on the host once per cycle."""
    from .linalg import IdentityOperator, make_linear_operator

    assert len(A.shape) == 2 and A.shape[0] == A.shape[1]
    op, sp, b, x, on_device = _setup(A, b, x0, "linalg.gmres")
    n = op.shape[0]
    M = IdentityOperator(op.shape, dtype=op.dtype) if M is None else make_linear_operator(M)
    b_norm = float(sp.nrm2(b)[0])
    if b_norm == 0:
        return _finish(b, on_device), 0
    atol = tol * b_norm if atol is None else max(float(atol), tol * b_norm)
    maxiter = 10 * _global_rows(op, b) if maxiter is None else maxiter
    restart = min(20 if restart is None else restart, _global_rows(op, b))
    if callback_type is None:
        callback_type = "pr_norm"
    if callback_type not in ("x", "pr_norm"):
        raise ValueError("Unknown callback_type: {}".format(callback_type))
    if callback is None:
        callback_type = None

    V = torch.empty((restart, n), dtype=b.dtype, device=b.device)
    H = torch.zeros((restart + 1, restart), dtype=b.dtype, device=b.device)
    iters = 0
    while True:
        mx = M.matvec(x)
        r = _residual(op, b, mx)
        r_norm = float(sp.nrm2(r)[0])
        if callback_type == "x":
            callback(_finish(mx, on_device))
        elif callback_type == "pr_norm" and iters > 0:
            callback(r_norm / b_norm)
        if r_norm <= atol or iters >= maxiter:
            break
        V[0] = r / r_norm
        H.zero_()
        for j in range(restart):
            u = op.matvec(M.matvec(V[j]))
            u, h = _project_out(sp, V[: j + 1], u)
            H[: j + 1, j] = h
            unorm = sp.nrm2(u)
            H[j + 1, j] = unorm[0]
            if j + 1 < restart:
                V[j + 1] = u / unorm
        e = np.zeros(restart + 1, dtype=np.float64)
        e[0] = r_norm
        y = np.linalg.lstsq(to_host(H).astype(np.float64), e, rcond=None)[0]
        x = x + torch.from_numpy(y).to(device=b.device, dtype=b.dtype) @ V
        iters += restart
    info = iters if (iters >= maxiter and not r_norm <= atol) else 0
    return _finish(mx, on_device), info


def _sym_ortho(a, b):
    """Stable Givens rotation (c, s, r) with c*a + s*b = r, -s*a + c*b = 0 (scipy's SymOrtho; linalg.py:905-934)."""
    if b == 0:
        return np.sign(a), 0.0, abs(a)
    if a == 0:
        return 0.0, np.sign(b), abs(b)
    if abs(b) > abs(a):
        tau = a / b
        s = np.sign(b) / np.sqrt(1.0 + tau * tau)
        return s * tau, s, b / s
    tau = b / a
    c = np.sign(a) / np.sqrt(1.0 + tau * tau)
    return c, c * tau, a / c


def lsqr(A, b, damp=0.0, atol=1e-6, btol=1e-6, conlim=1e8, iter_lim=None, show=False, calc_var=False, x0=None):
    """Paige & Saunders' LSQR for min ||A x - b||^2 + damp^2 ||x - x0||^2, rectangular A allowed (reference
    linalg.py:937-1413, which follows scipy 1.8.1).  Golub-Kahan bidiagonalisation with A.matvec / A.rmatvec on the
    device, the QR of the bidiagonal by plane rotations on the host.  Returns scipy's 10-tuple."""
    from .linalg import make_linear_operator

    runtime.require_cuda("linalg.lsqr")
    op = make_linear_operator(A)
    sp = _Space(op)
    m, n = op.shape
    on_device = is_device_array(b) and b.is_cuda
    dt = _work_dtype(op, b, x0)
    tdt = torch_dtype(dt)
    u = to_device(b, dtype=dt, copy=True).reshape(-1)
    assert u.shape[0] == m
    iter_lim = 2 * n if iter_lim is None else iter_lim
    var = torch.zeros(n, dtype=tdt, device=u.device)
    eps = float(np.finfo(np.float64).eps)
    ctol = 1.0 / conlim if conlim > 0 else 0.0
    dampsq = damp * damp
    anorm = acond = ddnorm = res2 = xnorm = xxnorm = z = 0.0
    cs2, sn2 = -1.0, 0.0
    itn = istop = 0

    bnorm = float(sp.nrm2(u)[0])
    if x0 is None:
        x = torch.zeros(n, dtype=tdt, device=u.device)
        beta = bnorm
    else:
        x = to_device(x0, dtype=dt, copy=True).reshape(-1)
        u = u - op.matvec(x)
        beta = float(sp.nrm2(u)[0])
    if beta > 0:
        u = u / beta
        v = op.rmatvec(u)
        alfa = float(sp.nrm2(v)[0])
    else:
        v = x.clone()
        alfa = 0.0
    if alfa > 0:
        v = v / alfa
    w = v.clone()
    rhobar, phibar = alfa, beta
    rnorm = r1norm = r2norm = beta
    arnorm = alfa * beta
    if arnorm == 0:
        return _finish(x, on_device), istop, itn, r1norm, r2norm, anorm, acond, arnorm, xnorm, _finish(var, on_device)

    while itn < iter_lim:
        itn += 1
        # next step of the bidiagonalisation: beta u = A v - alfa u ; alfa v = A^T u - beta v
        u = op.matvec(v) - alfa * u
        beta = float(sp.nrm2(u)[0])
        if beta > 0:
            u = u / beta
            anorm = np.sqrt(anorm * anorm + alfa * alfa + beta * beta + dampsq)
            v = op.rmatvec(u) - beta * v
            alfa = float(sp.nrm2(v)[0])
            if alfa > 0:
                v = v / alfa
        # eliminate the damping parameter, then the sub-diagonal of the bidiagonal
        if damp > 0:
            rhobar1 = np.sqrt(rhobar * rhobar + dampsq)
            cs1, sn1 = rhobar / rhobar1, damp / rhobar1
            psi = sn1 * phibar
            phibar = cs1 * phibar
        else:
            rhobar1, psi = rhobar, 0.0
        cs, sn, rho = _sym_ortho(rhobar1, beta)
        theta = sn * alfa
        rhobar = -cs * alfa
        phi = cs * phibar
        phibar = sn * phibar
        tau = sn * phi
        # x and the search direction w
        dk = w / rho
        x = x + (phi / rho) * w
        w = v - (theta / rho) * w
        ddnorm += float(sp.dot(dk, dk)[0])
        if calc_var:
            var = var + dk * dk
        # norm estimates (rotation on the right removes the super-diagonal of the upper bidiagonal)
        delta = sn2 * rho
        gambar = -cs2 * rho
        rhs = phi - delta * z
        zbar = rhs / gambar
        xnorm = np.sqrt(xxnorm + zbar * zbar)
        gamma = np.sqrt(gambar * gambar + theta * theta)
        cs2, sn2 = gambar / gamma, theta / gamma
        z = rhs / gamma
        xxnorm += z * z
        acond = anorm * np.sqrt(ddnorm)
        res1 = phibar * phibar
        res2 += psi * psi
        rnorm = np.sqrt(res1 + res2)
        arnorm = alfa * abs(tau)
        if damp > 0:
            r1sq = rnorm * rnorm - dampsq * xxnorm
            r1norm = np.sqrt(abs(r1sq))
            if r1sq < 0:
                r1norm = -r1norm
        else:
            r1norm = rnorm
        r2norm = rnorm
        # stopping rules of the paper, first the ones that depend on machine precision
        test1 = rnorm / bnorm
        test2 = arnorm / (anorm * rnorm + eps)
        test3 = 1.0 / (acond + eps)
        rtol = btol + atol * anorm * xnorm / bnorm
        if itn >= iter_lim:
            istop = 7
        if 1 + test3 <= 1:
            istop = 6
        if 1 + test2 <= 1:
            istop = 5
        if 1 + test1 / (1 + anorm * xnorm / bnorm) <= 1:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
        if show:
            print(f"{itn:6d} {rnorm:10.3e} {arnorm:10.3e} {test1:8.1e} {test2:8.1e} {anorm:8.1e} {acond:8.1e}")
        if istop != 0:
            break
    return (_finish(x, on_device), istop, itn, float(r1norm), float(r2norm), float(anorm), float(acond),
            float(arnorm), float(xnorm), _finish(var, on_device))


def _lanczos(op, sp, V, u, alpha, beta, start, end):
    """Lanczos steps start..end-1 with full re-orthogonalisation against the rows of V (linalg.py:1416-1424)."""
    for i in range(start, end):
        u = op.matvec(V[i])
        alpha[i] = sp.dot(V[i], u)[0]
        u, _ = _project_out(sp, V[: i + 1], u)
        bnorm = sp.nrm2(u)
        beta[i] = bnorm[0]
        if i >= end - 1:
            break
        V[i + 1] = u / bnorm
    return u


def _ritz(alpha, beta, beta_k, k, which):
    """Eigen-decomposition of the (arrowhead +) tridiagonal projected matrix on the host; keeps the k wanted pairs."""
    a, bt = to_host(alpha).astype(np.float64), to_host(beta).astype(np.float64)
    t = np.diag(a) + np.diag(bt[:-1], 1) + np.diag(bt[:-1], -1)
    if beta_k is not None:
        t[k, :k] = beta_k
        t[:k, k] = beta_k
    w, s = np.linalg.eigh(t)
    order = np.argsort(w) if which == "LA" else np.argsort(np.abs(w))
    keep = order[-k:]
    return w[keep], s[:, keep], float(bt[-1])


def eigsh(a, k=6, *, which="LM", ncv=None, maxiter=None, tol=0, return_eigenvectors=True):
    """k extremal eigenpairs of a real symmetric matrix by thick-restart Lanczos (reference linalg.py:1450-1569,
    after CuPy).  `which`: 'LM' largest magnitude, 'LA' largest algebraic.  Returns (w, x) ascending in w, x as a
    host (n, k) array; `w` only when `return_eigenvectors` is False."""
    from .linalg import make_linear_operator

    runtime.require_cuda("linalg.eigsh")
    if len(a.shape) != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("expected square matrix (shape: {})".format(a.shape))
    op = make_linear_operator(a)
    sp = _Space(op)
    n = op.shape[0]                                            # length of this process's vectors
    n_glob = int(getattr(op, "global_shape", op.shape)[0])     # size of the eigenproblem (row-sharded operators)
    dt = np.dtype(a.dtype)
    if dt.char not in "fd":
        raise TypeError("unsupprted dtype (actual: {})".format(a.dtype))
    if k <= 0:
        raise ValueError("k must be greater than 0 (actual: {})".format(k))
    if k >= n_glob:
        raise ValueError("k must be smaller than n (actual: {})".format(k))
    if which not in ("LM", "LA"):
        raise ValueError("which must be 'LM' or 'LA' (actual: {})".format(which))
    ncv = min(max(2 * k, k + 32), n_glob - 1) if ncv is None else min(max(ncv, k + 2), n_glob - 1)
    maxiter = 10 * n_glob if maxiter is None else maxiter
    if tol == 0:
        tol = float(np.finfo(dt).eps)
    tdt = torch_dtype(dt)
    dev = runtime.device
    alpha = torch.zeros(ncv, dtype=tdt, device=dev)
    beta = torch.zeros(ncv, dtype=tdt, device=dev)
    V = torch.empty((ncv, n), dtype=tdt, device=dev)
    u = to_device(np.random.random(n).astype(dt))
    V[0] = u / sp.nrm2(u)
    u = _lanczos(op, sp, V, u, alpha, beta, 0, ncv)
    spent = ncv
    w, s, beta_last = _ritz(alpha, beta, None, k, which)
    s_dev = torch.from_numpy(np.ascontiguousarray(s.T)).to(device=dev, dtype=tdt)    # (k, ncv)
    x = s_dev @ V                                                                      # rows = Ritz vectors
    beta_k = beta_last * s[-1, :]
    res = float(np.linalg.norm(beta_k))
    while res > tol and spent < maxiter:
        # thick restart: keep the k Ritz pairs, continue the recurrence from the last residual direction
        beta[:k] = 0
        alpha[:k] = torch.from_numpy(w).to(device=dev, dtype=tdt)
        V[:k] = x
        basis = V[:k]
        u, _ = _project_out(sp, basis, u)
        V[k] = u / sp.nrm2(u)
        u = op.matvec(V[k])
        alpha[k] = sp.dot(V[k], u)[0]
        u = u - alpha[k] * V[k]
        u = u - torch.from_numpy(beta_k).to(device=dev, dtype=tdt) @ basis
        bnorm = sp.nrm2(u)
        beta[k] = bnorm[0]
        V[k + 1] = u / bnorm
        u = _lanczos(op, sp, V, u, alpha, beta, k + 1, ncv)
        spent += ncv - k
        w, s, beta_last = _ritz(alpha, beta, beta_k, k, which)
        s_dev = torch.from_numpy(np.ascontiguousarray(s.T)).to(device=dev, dtype=tdt)
        x = s_dev @ V
        beta_k = beta_last * s[-1, :]
        res = float(np.linalg.norm(beta_k))
    order = np.argsort(w)
    if return_eigenvectors:
        return w[order].astype(dt), to_host(x).T[:, order]
    return w[order].astype(dt)
