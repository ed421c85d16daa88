"""The sharded path driven through the C ABI alone (tests/c/sharded_spmv_test.cu): NCCL communicator owned by
libb200sparse.so (b2s_comm_*), one host thread per GPU, no torch.distributed.  nranks = 1 runs on any GPU box."""
import os
import shutil
import subprocess

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

SRC = os.path.join(ROOT, "tests", "c", "sharded_spmv_test.cu")
BIN = os.path.join(ROOT, "tests", "c", "sharded_spmv_test")
LIBDIR = os.path.join(ROOT, "legate", "sparse_b200")


def _binary():
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(SRC):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
               "-o", BIN, SRC, "-I", os.path.join(ROOT, "include"), "-L", LIBDIR, "-lb200sparse",
               "-Xlinker", f"-rpath={LIBDIR}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    return BIN


@pytest.mark.parametrize("nranks", [1, 2])
def test_c_abi_sharded_spmv(nranks):
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    import legate.sparse_b200  # noqa: F401  (builds the library if needed)

    r = subprocess.run([_binary(), str(nranks)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"SHARDED_C_ABI_OK nranks={nranks}" in r.stdout
