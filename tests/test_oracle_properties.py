"""Size-independent properties of the CPU oracle on randomly drawn CSR matrices (hypothesis): the same properties
the GPU suite checks at the BASELINE sizes where the oracle itself is too slow to be the checker (linearity of
SpMV / SpMM, SpMM columns = SpMVs, (A B) x = A (B x), row-block shards reassemble the whole)."""
import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings
from hypothesis import strategies as st


@st.composite
def csr_matrices(draw, max_dim=40):
    m = draw(st.integers(1, max_dim))
    n = draw(st.integers(1, max_dim))
    density = draw(st.floats(0.0, 0.6))
    seed = draw(st.integers(0, 2**31 - 1))
    S = sp.random(m, n, density=density, format="csr", random_state=seed, dtype=np.float64)
    S.sort_indices()
    return S, seed


@settings(max_examples=60, deadline=None)
@given(csr_matrices(), st.floats(-4, 4), st.floats(-4, 4))
def test_spmv_is_linear_and_matches_scipy(oracle, As, a, b):
    S, seed = As
    rng = np.random.default_rng(seed)
    x, z = rng.standard_normal(S.shape[1]), rng.standard_normal(S.shape[1])
    f = lambda v: oracle.spmv(S.indptr, S.indices, S.data, v)
    assert np.array_equal(f(x), S @ x)                                   # same accumulation order as scipy
    assert np.allclose(f(a * x + b * z), a * f(x) + b * f(z), rtol=1e-12, atol=1e-12)


@settings(max_examples=40, deadline=None)
@given(csr_matrices(), st.integers(1, 9))
def test_spmm_columns_are_spmvs(oracle, As, k):
    S, seed = As
    X = np.random.default_rng(seed).standard_normal((S.shape[1], k))
    Y = oracle.spmm(S.indptr, S.indices, S.data, X)
    assert Y.shape == (S.shape[0], k)
    for j in range(k):
        assert np.array_equal(Y[:, j], oracle.spmv(S.indptr, S.indices, S.data, np.ascontiguousarray(X[:, j])))


@settings(max_examples=40, deadline=None)
@given(csr_matrices(25), st.integers(1, 25), st.integers(0, 2**31 - 1))
def test_spgemm_then_spmv_is_spmv_twice(oracle, As, p, seed2):
    A, seed = As
    B = sp.random(A.shape[1], p, density=0.3, format="csr", random_state=seed2, dtype=np.float64)
    B.sort_indices()
    cp, ci, cv = oracle.spgemm((A.indptr, A.indices, A.data), (B.indptr, B.indices, B.data), A.shape, B.shape,
                               sort_rows=True)
    assert cp[-1] == ci.shape[0] == cv.shape[0] and np.all(np.diff(cp) >= 0)
    for r in range(A.shape[0]):                                           # sorted, no duplicate columns
        seg = ci[cp[r] : cp[r + 1]]
        assert np.all(np.diff(seg) > 0)
    x = np.random.default_rng(seed).standard_normal(p)
    lhs = oracle.spmv(cp, ci, cv, x)
    rhs = oracle.spmv(A.indptr, A.indices, A.data, oracle.spmv(B.indptr, B.indices, B.data, x))
    assert np.allclose(lhs, rhs, rtol=1e-10, atol=1e-12)


@settings(max_examples=40, deadline=None)
@given(csr_matrices(), st.integers(1, 6))
def test_row_blocks_reassemble(oracle, As, nranks):
    S, seed = As
    x = np.random.default_rng(seed).standard_normal(S.shape[1])
    y = oracle.spmv(S.indptr, S.indices, S.data, x)
    parts = []
    covered = 0
    for r in range(nranks):
        lo, hi, klo, khi = oracle.row_block(S.indptr, r, nranks)
        assert lo == covered or lo == hi
        assert (klo, khi) == (S.indptr[lo], S.indptr[hi])
        covered = max(covered, hi)
        ip = S.indptr[lo : hi + 1] - S.indptr[lo]
        seg = slice(klo, khi)
        parts.append(oracle.spmv(ip, S.indices[seg], S.data[seg], x))
    assert covered == S.shape[0]
    assert np.array_equal(np.concatenate(parts), y)
