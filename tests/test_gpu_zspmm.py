"""GPU parity tests for CSR SpMM (A @ dense, dense @ A): CUDA path through the C ABI vs the CPU oracle, the
golden scipy vectors and scipy live.  Mirrors reference tests/integration/test_csr_spmm.py:28-79 and adds
the kernel's shape edge cases (every k around the lane-group and column-chunk boundaries, strided operands,
64-bit indices, empty rows).  Tolerance: fp64 1e-12, fp32 1e-5 (relative to |A||X|)."""
import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp
import torch

from conftest import MTX_FILES, mtx_path

import legate.sparse_b200 as sparse
from legate.sparse_b200 import _ops

pytestmark = pytest.mark.gpu

TYPES = [np.float32, np.float64]


@pytest.fixture(params=[2, 1, 3, 4], ids=["tile", "row", "window", "tma-window"])
def kernel(request):
    """All SpMM kernels forced in turn (the default picks one per launch): the staged tile kernel with global X
    gathers, the one-row-per-group kernel, the staged tile kernel with a synchronously loaded X window, and the
    persistent TMA-staged X-window kernel (both window kernels fall back to gathers, tile by tile, when the columns
    of a tile span more rows than a stage holds; the TMA one hands ineligible operand shapes to the gather kernels)."""
    from legate.sparse_b200 import _lib
    assert _lib.lib.b2s_spmm_set_kernel(request.param) == 0
    yield request.param
    _lib.lib.b2s_spmm_set_kernel(0)


def _check(Y, A_sp, X, dt):
    ref = A_sp.astype(np.float64) @ np.asarray(X, dtype=np.float64)
    bound = abs(A_sp.astype(np.float64)) @ np.abs(np.asarray(X, dtype=np.float64))
    eps = 2e-6 if np.dtype(dt) == np.float32 else 1e-13
    err = np.abs(np.asarray(Y, dtype=np.float64) - ref)
    assert np.all(err <= eps * 64 * (bound + 1e-300) + 1e-300), float(err.max())


@pytest.mark.parametrize("filename", MTX_FILES)
@pytest.mark.parametrize("b_type", TYPES)
@pytest.mark.parametrize("c_type", TYPES)
def test_csr_spmm(filename, b_type, c_type):
    arr = sparse.io.mmread(mtx_path(filename)).tocsr().astype(b_type)
    s = sio.mmread(mtx_path(filename), spmatrix=False).tocsr().astype(b_type)
    c = np.asarray(arr.todense()).astype(c_type)
    res = arr @ c
    assert res.dtype == np.result_type(b_type, c_type)
    assert np.allclose(res, s @ c)
    result = np.zeros(arr.shape, dtype=np.result_type(b_type, c_type))
    arr.dot(c, out=result)
    assert np.allclose(result, s @ c)
    with pytest.raises(ValueError):
        arr.dot(c, out=np.zeros(arr.shape, dtype=np.int64))


@pytest.mark.parametrize("filename", MTX_FILES)
@pytest.mark.parametrize("idim", [2, 4, 8, 16])
def test_csr_spmm_rmatmul(filename, idim):
    arr = sparse.io.mmread(mtx_path(filename)).tocsr()
    s = sio.mmread(mtx_path(filename), spmatrix=False).tocsr()
    x = np.ones((idim, arr.shape[0]))
    assert np.allclose(arr.__rmatmul__(x), x @ s)
    assert np.allclose(x @ arr, x @ s)  # numpy defers to __rmatmul__ (__array_priority__)


@pytest.mark.parametrize("filename", MTX_FILES)
@pytest.mark.parametrize("b_type", TYPES)
@pytest.mark.parametrize("c_type", TYPES)
def test_csr_spmm_rmatmul_types(filename, b_type, c_type):
    arr = sparse.io.mmread(mtx_path(filename)).tocsr().astype(b_type)
    s = sio.mmread(mtx_path(filename), spmatrix=False).tocsr().astype(b_type)
    x = np.ones((8, arr.shape[0])).astype(c_type)
    assert np.allclose(arr.__rmatmul__(x), x @ s)


@pytest.mark.parametrize("key", [n.split(".")[0] for n in MTX_FILES])
def test_golden_vectors(golden, oracle, key):
    A = sparse.csr_array((golden[f"{key}_data"], golden[f"{key}_indices"], golden[f"{key}_indptr"]),
                         shape=(len(golden[f"{key}_indptr"]) - 1, golden[f"{key}_spmm_x"].shape[0]))
    Y = A @ golden[f"{key}_spmm_x"]
    assert np.allclose(Y, golden[f"{key}_spmm_y"], rtol=1e-12, atol=1e-12 * np.abs(golden[f"{key}_spmm_y"]).max())
    Yo = oracle.spmm(golden[f"{key}_indptr"], golden[f"{key}_indices"], golden[f"{key}_data"], golden[f"{key}_spmm_x"])
    assert np.allclose(Y, Yo, rtol=1e-12, atol=1e-12 * np.abs(Yo).max())


# k sweeps the lane-group sizes (1..32 packs), the vector/scalar split (k % VEC) and the multi-pass column chunks
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 12, 16, 17, 31, 32, 33, 48, 64, 65, 100, 128, 129, 256, 300, 520])
@pytest.mark.parametrize("dt", TYPES)
def test_k_sweep_vs_oracle(oracle, kernel, k, dt):
    rng = np.random.default_rng(k)
    m, n = 777, 513
    A_sp = sp.random(m, n, density=0.02, format="csr", random_state=k, dtype=np.float64).astype(dt)
    A_sp.sort_indices()
    X = rng.standard_normal((n, k)).astype(dt)
    A = sparse.csr_array(A_sp)
    Y = A @ X
    assert Y.shape == (m, k) and Y.dtype == dt
    _check(Y, A_sp, X, dt)
    Yo = oracle.spmm(A_sp.indptr, A_sp.indices, A_sp.data, X)
    _check(Yo, A_sp, X, dt)
    # column j of the SpMM equals the SpMV with column j
    j = k // 2
    y = A @ np.ascontiguousarray(X[:, j])
    tol = dict(rtol=1e-4, atol=1e-4) if dt == np.float32 else dict(rtol=1e-11, atol=1e-11)
    assert np.allclose(Y[:, j], y, **tol)


@pytest.mark.parametrize("dt", TYPES)
def test_row_shapes(kernel, dt):
    """Empty rows, one row longer than the staging buffer (direct path of the tile kernel), a tile that just fits,
    empty matrix, single column."""
    rng = np.random.default_rng(5)
    m, n, k = 300, 4000, 32
    rows = [np.array([], dtype=np.int64)] * m
    rows[0] = np.arange(n)                      # dense row
    rows[7] = np.array([3])
    rows[8] = np.array([1, 2, 3])
    rows[9] = np.array([0, 5, 6, 7, 4000 - 1])
    rows[299] = np.arange(0, n, 7)
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    indices = np.concatenate(rows)
    data = rng.standard_normal(indices.shape[0]).astype(dt)
    A_sp = sp.csr_array((data, indices, indptr), shape=(m, n))
    X = rng.standard_normal((n, k)).astype(dt)
    Y = sparse.csr_array(A_sp) @ X
    _check(Y, A_sp, X, dt)
    assert np.all(Y[1:7] == 0)
    E = sparse.csr_array(sp.csr_array((m, n), dtype=dt))
    assert np.all((E @ X) == 0) and (E @ X).shape == (m, k)
    assert (sparse.csr_array(A_sp) @ X[:, :1]).shape == (m, 1)


@pytest.mark.parametrize("dt", TYPES)
def test_device_operands_strided_and_out(kernel, dt):
    """torch operands: padded leading dimensions (ldx, ldy > k) and out= written in place."""
    rng = np.random.default_rng(11)
    m, n, k = 1000, 900, 24
    A_sp = sp.random(m, n, density=0.01, format="csr", random_state=3, dtype=np.float64).astype(dt)
    A = sparse.csr_array(A_sp)
    Xh = rng.standard_normal((n, 40)).astype(dt)
    Xd = torch.from_numpy(Xh).cuda()
    Yd = torch.full((m, 64), 7.0, dtype=Xd.dtype, device="cuda")
    for lo in (0, 1, 4):          # lo = 1 breaks the 16-byte alignment -> scalar path
        Xv, Yv = Xd[:, lo:lo + k], Yd[:, lo:lo + k]
        _ops.spmm(A._indptr, A._indices, A._data, Xv, Yv, A.shape)
        torch.cuda.synchronize()
        _check(Yv.cpu().numpy(), A_sp, Xh[:, lo:lo + k], dt)
        assert torch.all(Yd[:, lo + k:] == 7.0) and torch.all(Yd[:, :lo] == 7.0)   # padding untouched
        Yd.fill_(7.0)
    out = torch.empty((m, k), dtype=Xd.dtype, device="cuda")
    res = A.dot(Xd[:, :k].contiguous(), out=out)
    assert res is out
    _check(out.cpu().numpy(), A_sp, Xh[:, :k], dt)
    res2 = A @ Xd[:, :k]                     # non-contiguous view -> copied, device result
    assert isinstance(res2, torch.Tensor) and res2.is_cuda
    _check(res2.cpu().numpy(), A_sp, Xh[:, :k], dt)


@pytest.mark.parametrize("dt", TYPES)
@pytest.mark.parametrize("k", [4, 32])
def test_tiles_around_the_staging_capacity(kernel, dt, k):
    """Row lengths chosen so that consecutive CTA tiles fall below, at and above the staging capacity."""
    rng = np.random.default_rng(8)
    m, n = 1500, 5000
    lens = np.concatenate([np.full(300, 5), np.full(300, 47), np.full(300, 48), np.full(300, 49),
                           rng.integers(0, 200, 300)])
    indptr = np.concatenate([[0], np.cumsum(lens)])
    indices = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens])
    data = rng.standard_normal(indices.shape[0]).astype(dt)
    A_sp = sp.csr_array((data, indices, indptr), shape=(m, n))
    X = rng.standard_normal((n, k)).astype(dt)
    _check(sparse.csr_array(A_sp) @ X, A_sp, X, dt)


def test_wide_indices(monkeypatch):
    monkeypatch.setenv("B2S_INDEX_WIDTH", "64")
    rng = np.random.default_rng(2)
    A_sp = sp.random(500, 400, density=0.03, format="csr", random_state=9, dtype=np.float64)
    A = sparse.csr_array(A_sp)
    assert A.indices.dtype == np.int64 or str(A.indices.dtype).endswith("int64")
    X = rng.standard_normal((400, 32))
    _check(A @ X, A_sp, X, np.float64)


def test_bad_arguments():
    from legate.sparse_b200 import _lib
    L = _lib.lib
    rc = L.b2s_spmm_csr(5, 0, 0, 1, 1, 0, 1, None, None, None, None, 1, None, 1, None)
    assert rc != 0 and b"value type" in L.b2s_last_error()
    rc = L.b2s_spmm_csr(1, 0, 0, 4, 4, 0, 8, None, None, None, None, 4, None, 8, None)
    assert rc != 0 and b"leading" in L.b2s_last_error()
    assert L.b2s_spmm_set_kernel(9) != 0 and b"unknown SpMM kernel" in L.b2s_last_error()


def test_laplacian_k32_matches_spmv_columns():
    """The microbenchmark shape (dot_microbenchmark.py -op spmm -k 32) at a small size."""
    from legate.sparse_b200 import gallery
    A = gallery.laplacian_5pt(130, 130)
    n = A.shape[0]
    X = torch.rand((n, 32), dtype=torch.float64, device="cuda")
    Y = A @ X
    for j in (0, 13, 31):
        y = A @ X[:, j].contiguous()
        assert torch.allclose(Y[:, j], y, rtol=1e-12, atol=1e-9)


# ---- complex operands (the reference dispatches SpMV / SpMM over complex64/128 as well) -------------------
CTYPES = [np.complex64, np.complex128]


def _crand(rng, shape, dt):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)


@pytest.mark.parametrize("a_type", TYPES + CTYPES)
@pytest.mark.parametrize("x_type", TYPES + CTYPES)
def test_complex_spmv_and_spmm(a_type, x_type):
    if np.dtype(a_type).kind != "c" and np.dtype(x_type).kind != "c":
        pytest.skip("real x real is covered above")
    rng = np.random.default_rng(17)
    m, n = 333, 257
    S = sp.random(m, n, density=0.05, format="csr", random_state=4, dtype=np.float64)
    S.sort_indices()
    data = _crand(rng, S.nnz, a_type) if np.dtype(a_type).kind == "c" else rng.standard_normal(S.nnz).astype(a_type)
    S = sp.csr_array((data, S.indices, S.indptr), shape=(m, n))
    A = sparse.csr_array(S)
    assert A.dtype == np.dtype(a_type)
    common = np.result_type(a_type, x_type)
    tol = dict(rtol=2e-4, atol=2e-4) if common == np.complex64 else dict(rtol=1e-11, atol=1e-11)
    x = _crand(rng, n, x_type) if np.dtype(x_type).kind == "c" else rng.standard_normal(n).astype(x_type)
    y = A @ x
    assert y.dtype == common and y.shape == (m,)
    assert np.allclose(y, S @ x, **tol)
    assert np.allclose(A.dot(x.reshape(n, 1)), (S @ x).reshape(m, 1), **tol)
    X = _crand(rng, (n, 5), x_type) if np.dtype(x_type).kind == "c" else rng.standard_normal((n, 5)).astype(x_type)
    Y = A @ X
    assert Y.dtype == common and Y.shape == (m, 5)
    assert np.allclose(Y, S @ X, **tol)
    out = np.zeros(m, dtype=common)
    assert A.dot(x, out=out) is out and np.allclose(out, S @ x, **tol)
    with pytest.raises(ValueError):
        A.dot(x, out=np.zeros(m, dtype=np.float64))
    xd = torch.from_numpy(x).cuda()
    yd = A @ xd
    assert isinstance(yd, torch.Tensor) and yd.is_cuda and np.allclose(yd.cpu().numpy(), S @ x, **tol)


def test_complex_matrix_helpers():
    rng = np.random.default_rng(3)
    S = sp.random(40, 30, density=0.2, format="csr", random_state=1, dtype=np.float64)
    S = sp.csr_array((_crand(rng, S.nnz, np.complex128), S.indices, S.indptr), shape=S.shape)
    A = sparse.csr_array(S)
    x = _crand(rng, 40, np.complex128)
    assert np.allclose(A.conj().todense(), S.conj().toarray())
    assert np.allclose(A.T.conj() @ x, S.T.conj() @ x)
    E = A._real_expansion(np.float64)
    assert E.shape == (80, 60) and E.nnz == 4 * A.nnz and A._real_expansion(np.float64) is E
    Ed = np.asarray(E.todense())
    Sd = S.toarray()
    assert np.array_equal(Ed[0::2, 0::2], Sd.real) and np.array_equal(Ed[1::2, 1::2], Sd.real)
    assert np.array_equal(Ed[1::2, 0::2], Sd.imag) and np.array_equal(Ed[0::2, 1::2], -Sd.imag)
