"""SpGEMM C = A @ A timing on the BASELINE workloads: python tools/bench_spgemm.py [case ...]
cases: banded1m banded10m rmat16 rmat18 rmat20 rmat22  (append :scipy to also time scipy on the host)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import gallery  # noqa: E402


def run(name, A, with_scipy):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    C = A @ A
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    info = C.spgemm_info
    del C
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        C = A @ A
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        nnzc = C.nnz
        del C
    t = min(ts)
    sv = A.dtype.itemsize
    byts = A.nnz * (sv + 4) + info["products"] * (sv + 4) + nnzc * (sv + 4) + 3 * (A.shape[0] + 1) * 8
    line = (f"SPGEMM {name:12s} n={A.shape[0]} nnz(A)={A.nnz} products={info['products']} nnz(C)={nnzc} "
            f"dense_rows={info['dense_rows']} first={t_first*1e3:.1f} ms best={t*1e3:.1f} ms "
            f"{2*info['products']/t/1e9:.1f} GFLOP/s model {byts/t/1e9:.0f} GB/s")
    if with_scipy:
        S = A.to_scipy_sparse_csr()
        t0 = time.perf_counter()
        Cs = S @ S
        tsci = time.perf_counter() - t0
        line += f" | scipy {tsci*1e3:.0f} ms (x{tsci/t:.0f}) nnz {Cs.nnz}"
    print(line, flush=True)


for case in (sys.argv[1:] or ["banded1m:scipy", "rmat16:scipy", "rmat18"]):
    with_scipy = case.endswith(":scipy")
    c = case.split(":")[0]
    if c.startswith("banded"):
        n = {"banded1m": 1_000_000, "banded10m": 10_000_000}[c]
        A = gallery.banded(n, 11, np.float64)
    elif c.startswith("rmat"):
        A = gallery.rmat(int(c[4:]), 16, 42, np.float64)
    else:
        raise SystemExit(f"unknown case {c}")
    run(c, A, with_scipy)
    del A
    torch.cuda.empty_cache()
