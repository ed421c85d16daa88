"""GPU parity tests for CSR SpMV: CUDA path (through the C ABI) vs the CPU oracle, the golden scipy
vectors and scipy live.  Mirrors reference tests/integration/test_csr_dot.py:25-45 (files x matrix dtype
x vector dtype; 1-D, (n,1) and out= forms) and adds the kernel's structural edge cases.
Tolerance: fp64 1e-12 relative to the row's |A||x| (north star bar is 1e-6); fp32 1e-5."""
import zlib

import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp
import torch

from conftest import MTX_FILES, mtx_path

import legate.sparse_b200 as sparse
from legate.sparse_b200 import _lib, _ops

pytestmark = pytest.mark.gpu

TYPES = [np.float32, np.float64]


def _tol(dt):
    return dict(rtol=1e-5, atol=1e-5) if np.dtype(dt) == np.float32 else dict(rtol=1e-12, atol=1e-12)


def _close_rowscaled(y, ref, A_abs_x, dt):
    """|y - ref| <= eps_budget * (|A||x|)_i : a forward-error bound independent of cancellation."""
    eps = 2e-6 if np.dtype(dt) == np.float32 else 1e-13
    err = np.abs(np.asarray(y, dtype=np.float64) - np.asarray(ref, dtype=np.float64))
    return np.all(err <= eps * (A_abs_x + 1e-300) * 64 + 1e-300)


@pytest.mark.parametrize("filename", MTX_FILES)
@pytest.mark.parametrize("mat_type", TYPES)
@pytest.mark.parametrize("vec_type", TYPES)
@pytest.mark.parametrize("col_split", [True, False])
def test_csr_dot(filename, mat_type, vec_type, col_split):
    arr = sparse.io.mmread(mtx_path(filename)).tocsr().astype(mat_type)
    s = sio.mmread(mtx_path(filename), spmatrix=False).tocsr().astype(mat_type)
    rng = np.random.default_rng(0)
    vec = rng.random(arr.shape[0]).astype(vec_type)
    assert np.allclose(arr @ vec, s @ vec)
    assert np.allclose(arr.dot(vec, spmv_domain_part=col_split), s.dot(vec))
    out_type = np.result_type(mat_type, vec_type)
    result_l = np.zeros(arr.shape[0], dtype=out_type)
    arr.dot(vec, spmv_domain_part=col_split, out=result_l)
    assert np.allclose(result_l, s.dot(vec))
    vec = rng.random((arr.shape[0], 1)).astype(vec_type)
    assert np.allclose(arr @ vec, s @ vec)
    assert np.allclose(arr.dot(vec), s.dot(vec))
    result_l = np.zeros((arr.shape[0], 1), dtype=out_type)
    arr.dot(vec, out=result_l)
    assert np.allclose(result_l, s.dot(vec))
    assert (arr @ vec).shape == (arr.shape[0], 1)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_column_split_product(oracle, dt):
    """`spmv_domain_part=True` (reference csr.py:869-927, spmv.cu:125-153): x is partitioned into column blocks and the
    partial products are reduced into y.  Same result as the row-split kernel (up to the order of the partial sums)
    and as the oracle, for every block count, with empty blocks, empty rows, and after an in-place edit of .data."""
    rng = np.random.default_rng(31)
    nrows, ncols = 4001, 9000
    lens = rng.integers(0, 40, nrows)
    lens[100:200] = 0
    ip, ix, dv = _random_csr(rng, nrows, ncols, lens, dt)
    ix = np.where(np.arange(len(ix)) % 7 == 0, ix % 1000, ix)          # crowd some entries into the first block
    order = np.lexsort((ix, np.repeat(np.arange(nrows), lens)))          # keep rows sorted by column
    ix, dv = ix[order], dv[order]
    A = sparse.csr_array((dv, ix, ip), shape=(nrows, ncols))
    x = rng.random(ncols).astype(dt)
    ref = oracle.spmv(ip, ix, dv, x)
    absx = oracle.spmv(ip, ix, np.abs(dv), np.abs(x))
    xd = torch.from_numpy(x).cuda()
    y_row = (A @ xd).cpu().numpy()
    y_col = A.dot(xd, spmv_domain_part=True).cpu().numpy()
    assert _close_rowscaled(y_col, ref, absx, dt) and _close_rowscaled(y_col, y_row, absx, dt)
    assert np.array_equal(A.dot(x, spmv_domain_part=True), y_col)        # host vectors take the same path
    for nb in (1, 3, 9, 16):
        y = torch.full((nrows,), 7.0, dtype=xd.dtype, device="cuda")
        A._dot_col_split(xd, y, nblocks=nb)
        assert _close_rowscaled(y.cpu().numpy(), ref, absx, dt)
        assert len(A._col_split(nb)) <= nb
    # all columns in the first of 4 blocks: three blocks are empty and skipped
    B = sparse.csr_array((dv, ix % 2000, ip), shape=(nrows, ncols))
    assert len(B._col_split(4)) == 1
    yb = torch.empty(nrows, dtype=xd.dtype, device="cuda")
    B._dot_col_split(xd, yb, nblocks=4)
    assert _close_rowscaled(yb.cpu().numpy(), oracle.spmv(ip, ix % 2000, dv, x), absx, dt)
    # in-place edit of the values: the cached blocks are rebuilt
    A.data[:] = A.data * 2
    assert _close_rowscaled(A.dot(xd, spmv_domain_part=True).cpu().numpy(), 2 * ref, 2 * absx, dt)
    # the automatic choice: only scattered matrices whose x exceeds L2
    assert not A._wants_col_split(A._get_plan())


@pytest.mark.parametrize("key", [n.split(".")[0] for n in MTX_FILES])
def test_golden_vectors(golden, oracle, key):
    A = sparse.csr_array((golden[f"{key}_data"], golden[f"{key}_indices"], golden[f"{key}_indptr"]),
                         shape=(golden[f"{key}_indptr"].shape[0] - 1,) * 2)
    y = A @ golden[f"{key}_x"]
    assert np.allclose(y, golden[f"{key}_y"], rtol=1e-13, atol=1e-14)
    # short rows are reduced sequentially in reference order: expect bit-exact up to FMA contraction
    yo = oracle.spmv(golden[f"{key}_indptr"], golden[f"{key}_indices"], golden[f"{key}_data"], golden[f"{key}_x"])
    assert np.allclose(y, yo, rtol=1e-14, atol=1e-15)
    if key == "test":
        assert np.array_equal(A @ np.arange(1.0, 6.0), golden["test_kat_y"])


def test_wrong_out_dtype_and_shapes():
    A = sparse.io.mmread(mtx_path("cage4.mtx")).tocsr()
    with pytest.raises(ValueError, match="not consistent"):
        A.dot(np.ones(9, dtype=np.float64), out=np.zeros(9, dtype=np.float32))
    with pytest.raises(AssertionError):
        A.dot(np.ones(8))
    with pytest.raises(NotImplementedError):
        A.dot(np.ones((9, 2, 2)))


def test_device_resident_vectors():
    A = sparse.io.mmread(mtx_path("karate.mtx")).tocsr()
    s = A.to_scipy_sparse_csr()
    x = torch.rand(34, dtype=torch.float64, device="cuda")
    y = A @ x
    assert isinstance(y, torch.Tensor) and y.is_cuda
    assert np.allclose(y.cpu().numpy(), s @ x.cpu().numpy())
    out = torch.zeros(34, dtype=torch.float64, device="cuda")
    r = A.dot(x, out=out)
    assert r.data_ptr() == out.data_ptr()
    assert np.allclose(out.cpu().numpy(), s @ x.cpu().numpy())


def _random_csr(rng, nrows, ncols, row_lens, dtype):
    row_lens = np.asarray(row_lens, dtype=np.int64)
    indptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(row_lens, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = rng.integers(0, ncols, nnz)
    # sort within rows (not required by the kernel, but makes scipy happy); duplicates are fine (summed)
    rows = np.repeat(np.arange(nrows), row_lens)
    order = np.lexsort((indices, rows))
    indices = indices[order]
    data = rng.standard_normal(nnz).astype(dtype)
    return indptr, indices, data


STRUCTURES = {
    "uniform5": lambda rng, n: np.full(n, 5),
    "uniform32": lambda rng, n: np.full(n, 32),
    "uniform8": lambda rng, n: np.full(n, 8),
    "uniform6": lambda rng, n: np.full(n, 6),
    "uniform64": lambda rng, n: np.full(n, 64),
    "uniform200": lambda rng, n: np.full(n, 200),
    "ragged": lambda rng, n: rng.integers(0, 40, n),
    "mostly_empty": lambda rng, n: (rng.random(n) < 0.05) * rng.integers(1, 9, n),
    "all_empty": lambda rng, n: np.zeros(n, dtype=np.int64),
    "one_long_row": lambda rng, n: np.where(np.arange(n) == n // 3, 30000, 3),
    "long_rows": lambda rng, n: np.where(np.arange(n) % 97 == 0, 6000, 2),
    "leading_empty_then_long": lambda rng, n: np.where(np.arange(n) == n - 1, 20000, 0),
    "powerlaw": lambda rng, n: np.minimum((rng.pareto(1.2, n) * 4).astype(np.int64), 50000),
}


@pytest.mark.parametrize("structure", sorted(STRUCTURES))
@pytest.mark.parametrize("dtype", TYPES)
@pytest.mark.parametrize("wide", [False, True])
def test_structures_vs_oracle(oracle, structure, dtype, wide):
    """Every kernel branch: sequential / sub-warp / warp phase-B, chunk-crossing tails, empty tiles, empty
    rows, plus int32 and int64 index instantiations -- compared against the CPU oracle."""
    rng = np.random.default_rng(zlib.crc32(structure.encode()))
    nrows, ncols = 5003, 4099
    lens = STRUCTURES[structure](rng, nrows)
    indptr, indices, data = _random_csr(rng, nrows, ncols, lens, dtype)
    x = rng.standard_normal(ncols).astype(dtype)
    idt = torch.int64 if wide else torch.int32
    d_ptr = torch.from_numpy(indptr).to("cuda", idt)
    d_idx = torch.from_numpy(indices).to("cuda", idt)
    d_val = torch.from_numpy(data).cuda()
    d_x = torch.from_numpy(x).cuda()
    y = torch.full((nrows,), float("nan"), dtype=d_val.dtype, device="cuda")
    plan = _ops.spmv_plan(d_ptr, d_idx, (nrows, ncols), int(indptr[-1]), dtype)
    plan.set_kernel(False)  # exercise the staged-tile kernel even when the matrix is judged scattered
    _ops.spmv(d_ptr, d_idx, d_val, d_x, y, (nrows, ncols), plan=plan)
    ref = oracle.spmv(indptr, indices, data, x)
    absx = oracle.spmv(indptr, indices, np.abs(data).astype(np.float64), np.abs(x).astype(np.float64))
    got = y.cpu().numpy()
    assert not np.isnan(got).any()
    assert _close_rowscaled(got, ref, absx, dtype), np.abs(got - ref).max()
    if plan.tma:
        # every kernel flavour (generic / uniform-row registers / one lane per short row) on the same tiles, and the
        # accumulating form y += A x
        auto = 2 if plan.short_rows else (1 if plan.uniform else 0)
        for flavor in (0, 1, 2):
            plan.set_flavor(flavor)
            yf = torch.full_like(y, float("nan"))
            _ops.spmv(d_ptr, d_idx, d_val, d_x, yf, (nrows, ncols), plan=plan)
            assert _close_rowscaled(yf.cpu().numpy(), ref, absx, dtype), (flavor, structure)
            y0 = torch.from_numpy(rng.standard_normal(nrows).astype(dtype)).cuda()
            ya = y0.clone()
            _ops.spmv_add(d_ptr, d_idx, d_val, d_x, ya, (nrows, ncols), plan)
            assert _close_rowscaled((ya - y0).cpu().numpy(), ref, absx + np.abs(y0.cpu().numpy()), dtype), (flavor, "add")
        plan.set_flavor(auto)
    # plan-free kernel (plan=NULL) must agree too
    y2 = torch.full_like(y, float("nan"))
    _ops.spmv(d_ptr, d_idx, d_val, d_x, y2, (nrows, ncols), plan=None)
    assert _close_rowscaled(y2.cpu().numpy(), ref, absx, dtype)
    # fused dot epilogue: sum_i w_i y_i
    w = torch.from_numpy(rng.standard_normal(nrows).astype(dtype)).cuda()
    out = torch.zeros(1, dtype=d_val.dtype, device="cuda")
    y3 = torch.empty_like(y)
    _ops.spmv_dot(d_ptr, d_idx, d_val, d_x, y3, w, out, (nrows, ncols), plan)
    assert torch.equal(y3, y)
    ref_dot = float(np.dot(w.cpu().numpy().astype(np.float64), got.astype(np.float64)))
    scale = float(np.dot(np.abs(w.cpu().numpy()).astype(np.float64), absx)) + 1e-300
    assert abs(float(out[0]) - ref_dot) <= (1e-5 if dtype == np.float32 else 1e-12) * scale


def test_unaligned_base_pointers(oracle):
    """indices/vals views that start at an odd element offset take the scalar-load path."""
    rng = np.random.default_rng(11)
    nrows = ncols = 3000
    indptr, indices, data = _random_csr(rng, nrows, ncols, rng.integers(0, 20, nrows), np.float64)
    x = rng.standard_normal(ncols)
    pad_i = torch.zeros(indices.shape[0] + 1, dtype=torch.int32, device="cuda")
    pad_v = torch.zeros(data.shape[0] + 1, dtype=torch.float64, device="cuda")
    pad_i[1:] = torch.from_numpy(indices).to("cuda", torch.int32)
    pad_v[1:] = torch.from_numpy(data).cuda()
    d_idx, d_val = pad_i[1:], pad_v[1:]
    assert d_idx.data_ptr() % 16 != 0
    d_ptr = torch.from_numpy(indptr).to("cuda", torch.int32)
    y = torch.empty(nrows, dtype=torch.float64, device="cuda")
    plan = _ops.spmv_plan(d_ptr, d_idx, (nrows, ncols), int(indptr[-1]), np.float64)
    _ops.spmv(d_ptr, d_idx, d_val, torch.from_numpy(x).cuda(), y, (nrows, ncols), plan=plan)
    assert np.allclose(y.cpu().numpy(), oracle.spmv(indptr, indices, data, x), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("cfg", range(12))
@pytest.mark.parametrize("dtype", TYPES)
def test_all_tile_configs(oracle, cfg, dtype):
    rng = np.random.default_rng(100 + cfg)
    nrows = ncols = 20011
    lens = np.where(np.arange(nrows) % 501 == 0, 9000, rng.integers(0, 12, nrows))
    indptr, indices, data = _random_csr(rng, nrows, ncols, lens, dtype)
    x = rng.standard_normal(ncols).astype(dtype)
    assert _lib.lib.b2s_spmv_num_configs() == 14
    try:
        _lib.check(_lib.lib.b2s_spmv_set_config(cfg, cfg % 3))
        A = sparse.csr_array((data, indices, indptr), shape=(nrows, ncols))
        plan = A._get_plan()
        assert plan.config == cfg
        plan.set_kernel(False)
        y = A @ x
    finally:
        _lib.check(_lib.lib.b2s_spmv_set_config(-1, 0))
    ref = oracle.spmv(indptr, indices, data, x)
    absx = oracle.spmv(indptr, indices, np.abs(data).astype(np.float64), np.abs(x).astype(np.float64))
    assert _close_rowscaled(y, ref, absx, dtype)


def _laplacian_5pt(n1, n2, dtype=np.float64):
    """pde.py:124-163 operator on an n1 x n2 interior grid, assembled directly in CSR."""
    N = n1 * n2
    i = np.arange(N, dtype=np.int64)
    a = float((n1 + 1) ** 2)
    g = float((n2 + 1) ** 2)
    c = -2 * a - 2 * g
    cols = np.stack([i - n1, i - 1, i, i + 1, i + n1], axis=1)
    vals = np.tile(np.array([g, a, c, a, g], dtype=dtype), (N, 1))
    valid = np.ones((N, 5), dtype=bool)
    valid[:, 0] = i >= n1
    valid[:, 1] = (i % n1) != 0
    valid[:, 3] = (i % n1) != n1 - 1
    valid[:, 4] = i < N - n1
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=indptr[1:])
    return indptr, cols[valid], vals[valid], N


def test_large_laplacian_properties():
    """BASELINE config 2 at full size (N ~ 10M, nnz ~ 50M, fp64): size-independent properties.
    A*1 has a closed form (row sums), A is symmetric (x.Ay == y.Ax), linear, and must equal the
    plan-free kernel."""
    n1 = n2 = 3162
    indptr, indices, data, N = _laplacian_5pt(n1, n2)
    assert N == 9998244 and indptr[-1] == 5 * N - 4 * n1 == 49978572
    A = sparse.csr_array((data, indices, indptr), shape=(N, N))
    ones = torch.ones(N, dtype=torch.float64, device="cuda")
    y1 = (A @ ones).cpu().numpy()
    rowsum = np.add.reduceat(data, indptr[:-1])
    assert np.allclose(y1, rowsum, rtol=1e-13, atol=1e-6)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(N, dtype=torch.float64, device="cuda", generator=g)
    z = torch.rand(N, dtype=torch.float64, device="cuda", generator=g)
    Ax, Az = A @ x, A @ z
    lhs, rhs = float(torch.dot(z, Ax)), float(torch.dot(x, Az))
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
    Axz = A @ (2.0 * x - 3.0 * z)
    assert torch.allclose(Axz, 2.0 * Ax - 3.0 * Az, rtol=1e-11, atol=1e-4)
    y2 = torch.empty_like(Ax)
    _ops.spmv(A.indptr, A.indices, A.data, x, y2, A.shape, plan=None)
    assert torch.allclose(Ax, y2, rtol=1e-13, atol=1e-7)
    # spot-check 1000 random rows against a direct numpy evaluation
    rows = np.random.default_rng(1).integers(0, N, 1000)
    xh = x.cpu().numpy()
    for r in rows[:1000]:
        lo, hi = indptr[r], indptr[r + 1]
        assert abs(float(Ax[r]) - float(np.dot(data[lo:hi], xh[indices[lo:hi]]))) <= 1e-8 * abs(data[lo:hi]).max()


def test_plan_kernel_choice_by_column_locality():
    """Stencil / banded matrices go to the TMA-staged tile kernel, uniformly random columns to the
    row-group kernel (they are L1-tag-bound: 32 distinct x lines per warp-wide gather)."""
    from legate.sparse_b200 import gallery

    L5 = gallery.laplacian_5pt(300, 300, np.float64)
    B = gallery.banded(100000, 11, np.float64)
    R = gallery.random_fixed(100000, 100000, 32, np.float32)
    pl, pb, pr = L5._get_plan(), B._get_plan(), R._get_plan()
    assert not pl.scattered and pl.lines_per_warp < 12
    assert not pb.scattered and pb.lines_per_warp < 6
    assert not pl.uniform and not pb.uniform      # short odd rows: one lane per row, not the shuffle-tree path
    assert pl.short_rows and pb.short_rows and pl.tma
    assert gallery.banded(100000, 32, np.float32)._get_plan().uniform
    assert pr.scattered and pr.lines_per_warp > 28 and pr.config == 8   # deep-MLP tile shape for scattered fp32, x via ld.global.cg
    # scattered SHORT rows (a column block of a random shard): 8-warp tile shape, one lane per row, x via ld.global.cg
    rng = np.random.default_rng(9)
    ip, ix, dv = _random_csr(rng, 100000, 100000, rng.integers(1, 8, 100000), np.float32)
    R4 = sparse.csr_array((dv, ix, ip), shape=(100000, 100000))._get_plan()
    assert R4.scattered and R4.short_rows and R4.tma and R4.config == 13
    # ... and uniform short rows (exactly 4 per row = one 16-byte group per lane) take the register path
    U4 = gallery.random_fixed(100000, 100000, 4, np.float32)._get_plan()
    assert U4.scattered and U4.uniform and U4.tma and U4.config == 13
    # both kernel families give the same answer on the same plan
    x = torch.rand(100000, dtype=torch.float32, device="cuda")
    y1 = R @ x
    pr.set_kernel(True)   # plan-free row-group kernel on the same matrix
    y2 = R @ x
    assert torch.allclose(y1, y2, rtol=1e-4, atol=1e-4)


def test_plan_rejects_mismatched_matrix():
    from legate.sparse_b200 import gallery

    A = gallery.banded(1000, 5, np.float64)
    B = gallery.banded(2000, 5, np.float64)
    x = torch.rand(2000, dtype=torch.float64, device="cuda")
    y = torch.empty(2000, dtype=torch.float64, device="cuda")
    with pytest.raises(_lib.B200SparseError, match="different matrix"):
        _ops.spmv(B.indptr, B.indices, B.data, x, y, B.shape, plan=A._get_plan())


@pytest.mark.parametrize("filename", MTX_FILES)
def test_balance_row_partitions(filename):
    """reference tests/integration/test_csr_misc.py:26-37: SpMV still equals scipy after arr.balance()."""
    arr = sparse.io.mmread(mtx_path(filename)).tocsr()
    arr.balance()
    s = sio.mmread(mtx_path(filename), spmatrix=False).tocsr()
    vec = np.random.default_rng(2).random(arr.shape[0])
    assert np.allclose(arr @ vec, s @ vec)
    assert np.allclose(arr.dot(vec), s.dot(vec))
    vec = np.random.default_rng(3).random((arr.shape[0], 1))
    assert np.allclose(arr @ vec, s @ vec)


@pytest.mark.parametrize("filename", MTX_FILES)
def test_csr_transpose_and_diagonal(filename):
    """reference test_csr_misc.py:40-44 (transpose) + the diagonal kernel used by GMG."""
    arr = sparse.io.mmread(mtx_path(filename)).tocsr()
    s = sio.mmread(mtx_path(filename), spmatrix=False).tocsr()
    assert np.array_equal(arr.T.todense(), np.ascontiguousarray(s.T.toarray()))
    assert np.array_equal(arr.diagonal().cpu().numpy(), s.diagonal())


@pytest.mark.parametrize("kind", ["laplacian", "banded32", "random"])
@pytest.mark.parametrize("pinned", [True, False])
def test_host_vectors_pipelined_path(kind, pinned, monkeypatch):
    """numpy x / out take the chunk-pipelined H2D -> tiles -> D2H path; it must equal the device-resident
    product bit for bit (same tiles, same kernel) for stencil, ELL-like and scattered matrices."""
    from legate.sparse_b200 import gallery

    if kind == "laplacian":
        A = gallery.laplacian_5pt(400, 300, np.float64)
    elif kind == "banded32":
        A = gallery.banded(150000, 32, np.float32)
    else:
        A = gallery.random_fixed(60000, 60000, 32, np.float32)
    plan = A._get_plan()
    if kind != "random":
        assert len(plan.chunks) == 16
        wins = [(c[4], c[5]) for c in plan.chunks]
        assert all(b > a for a, b in wins) and wins[0][1] < A.shape[1] // 2   # real windows, not the whole vector
    n = A.shape[1]
    tdt = torch.float64 if A.dtype == np.float64 else torch.float32
    xh = torch.rand(n, dtype=tdt)
    yh = torch.empty(A.shape[0], dtype=tdt)
    if pinned:
        xh, yh = xh.pin_memory(), yh.pin_memory()
    x_np, y_np = xh.numpy(), yh.numpy()
    ref = (A @ xh.cuda()).cpu().numpy()
    got = A @ x_np
    assert isinstance(got, np.ndarray) and np.array_equal(got, ref)
    r = A.dot(x_np, out=y_np)
    assert r is y_np and np.array_equal(y_np, ref)
    monkeypatch.setenv("B2S_PIPE_DIRECT", "1")   # option: tiles store y straight into a pinned out (pageable: falls back)
    y_np[:] = 0
    A.dot(x_np, out=y_np)
    assert np.array_equal(y_np, ref)
    monkeypatch.delenv("B2S_PIPE_DIRECT")
    for stages in ("3", "16"):
        monkeypatch.setenv("B2S_PIPE_CHUNKS", stages)
        y_np[:] = 0
        A.dot(x_np, out=y_np)
        assert np.array_equal(y_np, ref)
    monkeypatch.delenv("B2S_PIPE_CHUNKS")
    for pat, align in (("1,3,4,4,3,1", "0"), ("4,4,4,4", "100"), ("16", "512")):
        monkeypatch.setenv("B2S_PIPE_PATTERN", pat)
        monkeypatch.setenv("B2S_PIPE_ALIGN", align)
        y_np[:] = 0
        A.dot(x_np, out=y_np)
        assert np.array_equal(y_np, ref)
    monkeypatch.delenv("B2S_PIPE_PATTERN")
    monkeypatch.delenv("B2S_PIPE_ALIGN")
    monkeypatch.setenv("B2S_PIPELINE", "0")
    assert np.array_equal(A @ x_np, ref)
    # (n, 1) host vectors
    monkeypatch.setenv("B2S_PIPELINE", "1")
    assert np.array_equal((A @ x_np.reshape(-1, 1)).reshape(-1), ref)


def test_plan_entries_follow_the_tile_rule():
    """White-box check of b2s_spmv_plan_create: tile t starts at the first row r with indptr[r] + r >= t*T,
    its entry holds that row, indptr[row] and the row-shape code of the tile (common row length L, or -longest when the
    rows differ but none exceeds 32, else 0); so every
    tile has <= T rows and all rows but the last fit in T + 4 nonzeros."""
    rng = np.random.default_rng(77)
    nrows = 30011
    lens = rng.integers(0, 9, nrows)
    lens[5000:9000] = 6            # a uniform stretch
    lens[12345] = 7000             # one row longer than several tiles
    lens[20000:20400] = 0          # a run of empty rows
    indptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    rows = np.repeat(np.arange(nrows), lens)
    indices = np.clip(rows + rng.integers(-40, 41, rows.shape[0]), 0, nrows - 1)   # columns near the diagonal
    data = rng.standard_normal(rows.shape[0])
    A = sparse.csr_array((data, indices, indptr), shape=(nrows, nrows))
    plan = A._get_plan()
    assert plan.config == 0 and not plan.scattered, (plan.config, plan.lines_per_warp)
    T = 2 * 128 * 4 - 4            # fp64 default tile shape: CAP = EPT(2) * 128 consumer threads * 4 groups
    ntiles = plan.tiles
    assert ntiles == -(-(nrows + int(indptr[-1])) // T)
    ent = plan.buf.cpu().numpy()[: 4 * (ntiles + 1)].reshape(-1, 4)
    k = ent[:, 0].astype(np.uint32).astype(np.int64) | (ent[:, 1].astype(np.int64) << 32)
    row, pad = ent[:, 2], ent[:, 3]
    s = indptr[:-1] + np.arange(nrows)            # start position of each row in the (rows + nnz) work list
    s = np.append(s, indptr[-1] + nrows)
    expect_row = np.searchsorted(s, np.arange(ntiles) * T, side="left")
    assert np.array_equal(row[:ntiles], expect_row) and row[ntiles] == nrows
    assert np.array_equal(k, indptr[row])
    assert (np.diff(row) <= T).all()
    for t in range(ntiles):
        r0, r1 = row[t], row[t + 1]
        if r1 > r0 + 1:
            assert indptr[r1 - 1] - indptr[r0] <= T        # all rows but the last fit in the staged chunk
        ls = lens[r0:r1]
        if r1 > r0 and ls[0] > 0 and (ls == ls[0]).all():
            want = int(ls[0])                                # uniform tile: the common row length
        elif r1 > r0 and ls.max() <= 32:
            want = -max(int(ls.max()), 1)                    # short rows: minus the longest
        else:
            want = 0
        assert pad[t] == want, (t, pad[t], want)
    assert (pad[:ntiles] == 6).sum() >= 20                 # the uniform stretch is recognised


def test_fused_entry_without_exchange_and_tile_ranges(oracle):
    """b2s_spmv_csr_fused with an empty exchange: explicit tile ranges in any order give the plain product, the fused
    dot works through it, and `accumulate` adds."""
    from legate.sparse_b200 import gallery

    A = gallery.laplacian_5pt(300, 200, np.float64)
    plan = A._get_plan()
    n = A.shape[0]
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    ref = A @ x
    nt = plan.tiles
    cuts = [(nt // 2, nt), (0, nt // 3), (nt // 3, nt // 2)]
    y = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    _ops.spmv_fused(A.indptr, A.indices, A.data, x, y, A.shape, plan, _ops.fuse_desc(cuts))
    assert torch.equal(y, ref)
    w = torch.rand(n, dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    y2 = torch.empty_like(y)
    _ops.spmv_fused(A.indptr, A.indices, A.data, x, y2, A.shape, plan, _ops.fuse_desc([(0, nt)]), w=w, dot_out=out)
    assert torch.equal(y2, ref)
    assert abs(float(out[0]) - float(torch.dot(w, ref))) <= 1e-12 * float(torch.dot(w.abs(), ref.abs()))
    y3 = torch.ones_like(y)
    _ops.spmv_fused(A.indptr, A.indices, A.data, x, y3, A.shape, plan, _ops.fuse_desc([(0, nt)], accumulate=True))
    assert torch.allclose(y3, ref + 1.0, rtol=1e-13, atol=1e-7)


def test_r32_full_size_vs_scipy_and_oracle(oracle):
    """BASELINE config 4 at full size (10M x 10M, 32 random entries per row, fp32): the device product against scipy
    (1e-6 of the row's |A||x| scale -- the north-star bar) and against the CPU oracle; fp64 on a 2M-row slice."""
    from legate.sparse_b200 import gallery

    n = 10_000_000
    A = gallery.random_fixed(n, n, 32, np.float32)
    assert A.nnz == 320_000_000
    plan = A._get_plan()
    assert plan.scattered and plan.uniform
    x = torch.rand(n, dtype=torch.float32, device="cuda")
    y = (A @ x).cpu().numpy()
    ip, ix, dv = A.indptr.cpu().numpy(), A.indices.cpu().numpy(), A.data.cpu().numpy()
    xh = x.cpu().numpy()
    S = sp.csr_array((dv, ix, ip), shape=A.shape)
    ref = S @ xh
    absx = sp.csr_array((np.abs(dv), ix, ip), shape=A.shape) @ xh
    assert np.all(np.abs(y - ref) <= 2e-6 * absx + 1e-30)
    rows = 1_000_000
    yo = oracle.spmv(ip[: rows + 1], ix[: 32 * rows], dv[: 32 * rows], xh, omp=True)
    assert np.all(np.abs(y[:rows] - yo) <= 2e-6 * absx[:rows] + 1e-30)
    del S, ref, absx
    # size-independent property at full size: linearity
    z = torch.rand(n, dtype=torch.float32, device="cuda")
    lhs = A @ (x + z)
    rhs = (A @ x) + (A @ z)
    scale = float(rhs.abs().max())
    assert float((lhs - rhs).abs().max()) <= 2e-5 * scale


def test_inplace_data_edits_invalidate_derived_matrices():
    """`.data` is a mutable device array (reference csr.py:264-287): after an in-place edit the cached promoted
    copy (f32 matrix x f64 vector), the cached transpose (dense @ A) and the complex expansion must follow."""
    S = sp.random(300, 300, density=0.05, random_state=np.random.default_rng(5), format="csr", dtype=np.float32)
    A = sparse.csr_array(S)
    x = np.random.default_rng(6).random(300)            # float64 operand -> promoted copy of A
    X = np.random.default_rng(7).random((4, 300))
    assert np.allclose(A @ x, S @ x) and np.allclose(X @ A, X @ S.toarray(), rtol=1e-5)
    A.data *= 2                                          # same tensor, same pointer, new values
    assert np.allclose(A @ x, 2 * (S @ x)), "stale promoted copy"
    assert np.allclose(X @ A, 2 * (X @ S.toarray()), rtol=1e-5), "stale transpose"
    A.data[:] = 1.0
    ones = sp.csr_array((np.ones_like(S.data), S.indices, S.indptr), shape=S.shape)
    assert np.allclose(A @ x, ones @ x)
    C = sparse.csr_array(sp.csr_array((S.data.astype(np.complex128) * (1 + 2j), S.indices, S.indptr), shape=S.shape))
    z = x + 1j * x[::-1]
    ref = C.to_scipy_sparse_csr() @ z
    assert np.allclose(C @ z, ref)
    C.data *= 1j
    assert np.allclose(C @ z, 1j * ref), "stale real expansion"
