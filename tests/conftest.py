import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MTX_FILES = ["test.mtx", "GlossGT.mtx", "Ragusa18.mtx", "cage4.mtx", "karate.mtx"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def mtx_path(name):
    return os.path.join(GOLDEN, name)


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (tests only)."""
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "golden.npz"))


def sample_spd(N, density, seed):
    """Seeded SPD test matrix, as reference tests/integration/utils/sample.py:25-44 +
    test_cg_solve.py:23-31: A = 0.5 (S + S^T) + N I with S = scipy.sparse.random(normal)."""
    import scipy.sparse as scpy
    import scipy.stats as stats

    class Normal(stats.rv_continuous):
        def _rvs(self, *args, size=None, random_state=None):
            return random_state.standard_normal(size)

    rv = Normal(seed=seed)()
    S = np.asarray(scpy.random(N, N, density=density, format="csr", dtype=np.float64, random_state=seed,
                               data_rvs=rv.rvs).todense())
    A = 0.5 * (S + S.T) + N * np.eye(N)
    xs = np.asarray(scpy.random(N, 1, density=density, format="csr", dtype=np.float64, random_state=seed,
                                data_rvs=Normal(seed=seed)().rvs).todense()).squeeze()
    return A, xs


def sample(N, D, density, seed):
    """Seeded random N x D matrix with normal entries and its dense test vector, as reference
    tests/integration/utils/sample.py:25-44 (`sample`, `sample_dense_vector`)."""
    import scipy.sparse as scpy
    import scipy.stats as stats

    class Normal(stats.rv_continuous):
        def _rvs(self, *args, size=None, random_state=None):
            return random_state.standard_normal(size)

    S = scpy.random(N, D, density=density, format="csr", dtype=np.float64, random_state=seed,
                    data_rvs=Normal(seed=seed)().rvs)
    xs = np.asarray(scpy.random(D, 1, density=density, format="csr", dtype=np.float64, random_state=seed,
                                data_rvs=Normal(seed=seed)().rvs).todense()).squeeze()
    return S, xs
