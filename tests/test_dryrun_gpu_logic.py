"""Runs the newest `-m gpu` test files through tests/dryrun_cpu.py (every tensor poses as a CUDA tensor, the leaf
launchers are the CPU oracle) so that the GPU-less suite already catches Python-level breakage -- shapes, dtypes,
dispatch, return conventions -- in code that otherwise only executes on the B200 box.  Says nothing about kernels."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

FILES = ["test_gpu_zspmm.py", "test_gpu_zy_krylov.py", "test_gpu_zzassembly.py", "test_gpu_cg.py", "test_gpu_spmv.py"]


@pytest.mark.parametrize("name", FILES)
def test_gpu_test_file_passes_the_cpu_dry_run(name):
    extra = []
    if name == "test_gpu_zzassembly.py":
        extra = ["-k", "not spectral_norm"]        # that one launches the real example in a subprocess
    if name == "test_gpu_spmv.py":
        extra = ["-k", "column_split"]             # host logic of the column-split product (blocks, caches, empty blocks)
    if name == "test_gpu_cg.py":
        extra = ["-k", "not full_size"]            # BASELINE-size problems are for the GPU box (minutes of scipy here)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_cpu.py"), os.path.join(ROOT, "tests", name), *extra],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
