"""Iterations/s of the Krylov solvers next to cg on the pde.py operator (5-pt Laplacian, SPD):
python tools/bench_krylov.py [grid=2048] [iters=200].  Each solver runs `iters` iterations (tolerance unreachable),
timed with CUDA events after one short warm-up call; cg (fused loop) is printed beside them as the yardstick."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import gallery, linalg  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
A = gallery.laplacian_5pt(g, g, np.float64)
A = A * -1.0            # positive definite, like examples/pde.py solves it
b = torch.ones(A.shape[0], dtype=torch.float64, device="cuda")


def timed(name, fn):
    fn(10)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn(iters)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"KRYLOV {name:9s} n={A.shape[0]} {iters / (ms * 1e-3):9.1f} it/s  {ms / iters * 1e3:8.1f} us/it", flush=True)


timed("cg", lambda k: linalg.cg(A, b, tol=1e-300, maxiter=k))
timed("cgs", lambda k: linalg.cgs(A, b, tol=1e-300, maxiter=k))
timed("bicg", lambda k: linalg.bicg(A, b, tol=1e-300, maxiter=k))
timed("bicgstab", lambda k: linalg.bicgstab(A, b, tol=1e-300, maxiter=k))
timed("gmres(20)", lambda k: linalg.gmres(A, b, tol=1e-300, maxiter=k, restart=20))
timed("lsqr", lambda k: linalg.lsqr(A, b, atol=0, btol=0, conlim=0, iter_lim=k))
