"""Device-side COO -> CSR assembly (`coo_array` built from CUDA tensors; reference sparse/coo.py:233-347 sort-by-key
+ base.py:30-48 counts -> pos): must equal scipy's `coo.tocsr()` with sorted indices, for shuffled triplets, empty
rows and both index widths.  Mirrors reference tests/integration/test_coo.py / test_csr_from_coo conversions."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import legate.sparse_b200 as sparse
from legate.sparse_b200.coo import coo_array

pytestmark = pytest.mark.gpu


def _triplets(m, n, nnz, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    flat = rng.choice(m * n, size=nnz, replace=False)          # no duplicates (coo.py:73-76)
    rng.shuffle(flat)
    return (flat // n).astype(np.int64), (flat % n).astype(np.int64), rng.standard_normal(nnz).astype(dtype)


@pytest.mark.parametrize("shape,nnz", [((50, 40), 300), ((1000, 1000), 20000), ((7, 5000), 900), ((3000, 3), 1500)])
@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_device_triplets_to_csr(shape, nnz, idx, dt):
    m, n = shape
    r, c, v = _triplets(m, n, nnz, seed=m + n, dtype=dt)
    ref = sp.coo_array((v, (r, c)), shape=shape).tocsr()
    ref.sort_indices()
    dev = "cuda"
    coo = coo_array((torch.from_numpy(v).to(dev), (torch.from_numpy(r).to(dev).to(idx), torch.from_numpy(c).to(dev).to(idx))),
                    shape=shape)
    assert coo.nnz == nnz and coo.dtype == np.dtype(dt) and coo.shape == shape
    A = coo.tocsr()
    assert A._data.is_cuda and A.shape == shape and A.nnz == nnz
    assert np.array_equal(A._indptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(A._indices.cpu().numpy(), ref.indices)
    assert np.array_equal(A._data.cpu().numpy(), ref.data)
    # the constructor form of the reference, csr_array((data, (row, col)), shape=...)
    B = sparse.csr_array((torch.from_numpy(v).to(dev), (torch.from_numpy(r).to(dev), torch.from_numpy(c).to(dev))),
                         shape=shape)
    assert np.array_equal(B._indices.cpu().numpy(), ref.indices) and np.array_equal(B._data.cpu().numpy(), ref.data)
    x = np.random.default_rng(1).random(n).astype(dt)
    assert np.allclose(B @ x, ref @ x, rtol=1e-4 if dt == np.float32 else 1e-12, atol=1e-5 if dt == np.float32 else 1e-12)


def test_round_trip_and_host_views():
    S = sp.random(200, 150, density=0.05, format="csr", random_state=2, dtype=np.float64)
    S.sort_indices()
    A = sparse.csr_array(S)
    coo = A.tocoo()
    assert coo.nnz == S.nnz
    assert np.array_equal(coo.todense(), S.toarray())
    c = S.tocoo()
    assert np.array_equal(coo.row, c.row) and np.array_equal(coo.col, c.col) and np.array_equal(coo.data, c.data)
    back = coo.tocsr()
    assert np.array_equal(back._indptr.cpu().numpy(), S.indptr) and np.array_equal(back._indices.cpu().numpy(), S.indices)
    T = coo.T.tocsr()
    St = sp.csr_array(S.T)
    St.sort_indices()
    assert np.array_equal(T._indices.cpu().numpy(), St.indices) and np.array_equal(T._data.cpu().numpy(), St.data)
    assert coo.astype(np.float32).tocsr().dtype == np.float32
    # empty matrix and inferred shape
    z = torch.zeros(0, dtype=torch.int64, device="cuda")
    E = coo_array((torch.zeros(0, dtype=torch.float64, device="cuda"), (z, z)), shape=(5, 4)).tocsr()
    assert E.nnz == 0 and E._indptr.cpu().tolist() == [0] * 6
    r = torch.tensor([3, 0], device="cuda")
    c = torch.tensor([1, 6], device="cuda")
    assert coo_array((torch.ones(2, dtype=torch.float64, device="cuda"), (r, c))).shape == (4, 7)


def test_wide_indices(monkeypatch):
    monkeypatch.setenv("B2S_INDEX_WIDTH", "64")
    r, c, v = _triplets(300, 200, 1000, seed=9)
    A = coo_array((torch.from_numpy(v).cuda(), (torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda())), shape=(300, 200)).tocsr()
    assert A._indices.dtype == torch.int64 and A._indptr.dtype == torch.int64
    ref = sp.coo_array((v, (r, c)), shape=(300, 200)).tocsr()
    x = np.random.default_rng(0).random(200)
    assert np.allclose(A @ x, ref @ x)


def test_spectral_norm_example():
    """examples/spectral_norm.py (the reference's example of the same name) runs on the GPU and agrees with scipy."""
    import os
    import re
    import subprocess
    import sys

    from conftest import ROOT

    def run(*extra):
        r = subprocess.run([sys.executable, "spectral_norm.py", "-n", "100", *extra], capture_output=True, text=True,
                           timeout=300, cwd=os.path.join(ROOT, "examples"))
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        return float(re.search(r"estimate: ([0-9.]+)", r.stdout).group(1))

    assert abs(run() - run("--package", "scipy")) < 1e-8


@pytest.mark.parametrize("name", ["test.mtx", "cage4.mtx", "karate.mtx"])
def test_coo_products_go_through_csr(name):
    """reference tests/integration/test_coo.py:52-69 (coo @ dense, coo * scalar)."""
    import scipy.io as sio

    from conftest import mtx_path

    arr = sparse.io.mmread(mtx_path(name)).tocoo()
    s = sio.mmread(mtx_path(name), spmatrix=False).tocoo()
    assert np.allclose(arr @ arr.todense(), s @ s.toarray())
    assert np.allclose((arr * 3.0).todense(), (s * 3.0).toarray())
    x = np.random.default_rng(0).random(arr.shape[1])
    assert np.allclose(arr @ x, s @ x)
    assert np.array_equal(arr.T.todense(), s.T.toarray())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("wide", [False, True])
def test_library_assembly_kernels_row_shapes(dtype, wide, monkeypatch):
    """b2s_coo_to_csr / b2s_csr_transpose (csrc/convert.cu) on every row-sort path: empty rows, rows of <= 32 entries
    (warp network), rows of hundreds / thousands (shared-memory network) and one row beyond 4096 entries (in-place
    global network); int32 and int64 index instantiations.  Bit-exact against scipy's canonical CSR."""
    import scipy.sparse as sp
    import torch

    import legate.sparse_b200 as sparse

    if wide:
        monkeypatch.setenv("B2S_INDEX_WIDTH", "64")
    rng = np.random.default_rng(3)
    m, n = 3000, 9000
    lens = rng.integers(0, 20, m)
    lens[7] = 33; lens[100] = 700; lens[101] = 4096; lens[1500] = 6000; lens[2999] = 1
    lens[200:260] = 0
    rows = np.repeat(np.arange(m), lens)
    cols = np.concatenate([rng.choice(n, size=int(k), replace=False) for k in lens]) if lens.sum() else np.zeros(0, int)
    vals = rng.standard_normal(rows.shape[0]).astype(dtype)
    perm = rng.permutation(rows.shape[0])            # triplets in random order
    rows, cols, vals = rows[perm], cols[perm], vals[perm]
    ref = sp.coo_array((vals, (rows, cols)), shape=(m, n)).tocsr()
    ref.sort_indices()
    A = sparse.coo_array((torch.from_numpy(vals).cuda(), (torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda())),
                         shape=(m, n)).tocsr()
    assert A.indices.dtype == (torch.int64 if wide else torch.int32)
    assert np.array_equal(A.indptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(A.indices.cpu().numpy(), ref.indices)
    assert np.array_equal(A.data.cpu().numpy(), ref.data)
    T = A.transpose()
    rt = ref.T.tocsr()
    rt.sort_indices()
    assert T.shape == (n, m)
    assert np.array_equal(T.indptr.cpu().numpy(), rt.indptr)
    assert np.array_equal(T.indices.cpu().numpy(), rt.indices)
    assert np.array_equal(T.data.cpu().numpy(), rt.data)
    # out-of-range rows are reported, not written
    with pytest.raises(ValueError, match="outside"):
        sparse.coo_array((torch.ones(2, dtype=torch.float64, device="cuda"),
                          (torch.tensor([0, 5], device="cuda"), torch.tensor([0, 1], device="cuda"))), shape=(3, 3)).tocsr()
