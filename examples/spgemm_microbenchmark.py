#!/usr/bin/env python
"""CSR x CSR SpGEMM throughput on a banded matrix -- the workload of the reference's
examples/spgemm_microbenchmark.py (B = A.copy(); C = A @ B; one warm-up then `-i` timed products).

    python examples/spgemm_microbenchmark.py -n 1000000 -i 25 [--package scipy]
"""
import argparse

import numpy as np

from common import select_package

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, default=1_000_000)
ap.add_argument("-i", type=int, default=25, dest="iters")
ap.add_argument("-nnz-per-row", type=int, default=11, dest="k")
ap.add_argument("--package", default="b200")
args = ap.parse_args()

name, timer, xp, sparse, _, on_device = select_package()
offsets = [d - args.k // 2 for d in range(args.k)]
A = sparse.diags([1] * args.k, offsets, shape=(args.n, args.n), format="csr", dtype=np.float64)
B = A.copy()
C = A @ B  # warm-up
timer.start()
for _ in range(args.iters):
    C = A @ B
ms = timer.stop()
print(f"Iterations / sec: {args.iters / (ms / 1e3):.3f}")
print(f"[{name}] n={args.n} nnz(A)={A.nnz} nnz(C)={C.nnz} {ms / args.iters:.2f} ms per product")
