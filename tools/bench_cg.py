"""CG throughput on the pde.py operator (BASELINE config 3): python tools/bench_cg.py [nx] [max_iter]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import _ops, gallery, linalg  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n1 = nx - 2
A = gallery.laplacian_5pt(n1, n1, np.float64)
N = A.shape[0]
b = torch.ones(N, dtype=torch.float64, device="cuda")
_ = A.dot(torch.zeros(N, dtype=torch.float64, device="cuda"))
PEAK = 6583.5
for fused in ("1", "0"):
    os.environ["B2S_CG_FUSED"] = fused
    linalg.cg(A, b, tol=1e-10, maxiter=30)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    x, iters = linalg.cg(A, b, tol=1e-10, maxiter=max_iter)
    e.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = s.elapsed_time(e)
    byts = (A.nnz * 12 + N * 20 + (9 if fused == "1" else 16) * 8 * N)
    print(f"nx={nx} N={N} nnz={A.nnz} fused={fused}: {iters} iters in {ms:.2f} ms (wall {wall*1e3:.1f}) -> "
          f"{iters/(ms*1e-3):.1f} it/s ; per-iter {ms/iters*1e3:.1f} us ; model bytes/iter {byts/1e6:.0f} MB -> "
          f"{byts*iters/(ms*1e-3)/1e9:.0f} GB/s ({byts*iters/(ms*1e-3)/1e9/PEAK:.2f} of peak)")

# individual kernels
def timeit(f, n=50):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e-3

xv = torch.rand(N, dtype=torch.float64, device="cuda")
yv = torch.rand(N, dtype=torch.float64, device="cuda")
pv = torch.rand(N, dtype=torch.float64, device="cuda")
qv = torch.rand(N, dtype=torch.float64, device="cuda")
a = torch.tensor([1.5], dtype=torch.float64, device="cuda")
bb = torch.tensor([2.5], dtype=torch.float64, device="cuda")
out = torch.zeros(1, dtype=torch.float64, device="cuda")
plan = A._get_plan()
for name, f, nbytes in [
    ("spmv", lambda: _ops.spmv(A.indptr, A.indices, A.data, xv, yv, A.shape, plan=plan), A.nnz * 12 + N * 20),
    ("spmv_dot", lambda: _ops.spmv_dot(A.indptr, A.indices, A.data, xv, yv, xv, out, A.shape, plan), A.nnz * 12 + N * 28),
    ("axpby", lambda: _ops.axpby(yv, xv, a, bb), 24 * N),
    ("dot", lambda: _ops.dot(xv, yv, out), 16 * N),
    ("nrm2", lambda: _ops.nrm2(xv, out), 8 * N),
    ("cg_update_xr", lambda: _ops.cg_update_xr(xv, yv, pv, qv, a, bb, out), 48 * N),
    ("torch copy", lambda: yv.copy_(xv), 16 * N),
]:
    t = timeit(f)
    print(f"  {name:14s} {t*1e6:8.1f} us  {nbytes/t/1e9:7.0f} GB/s  ({nbytes/t/1e9/PEAK:.2f} of peak)")
