#!/usr/bin/env python
"""SpMV / SpMM throughput on a banded matrix -- the workload of the reference's examples/dot_microbenchmark.py
(diags([1]*k, centred offsets, shape (n, n), csr), x = ones, one warm-up then `-i` timed products; prints
"Iterations / sec").  The Summit logs in the reference were produced with `-nnz-per-row 11 -n 10000000 -i 100`.

    python examples/dot_microbenchmark.py -n 10000000 -i 100 [-op spmm -k 32] [--package scipy]
"""
import argparse

import numpy as np

from common import select_package

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, default=10_000_000)
ap.add_argument("-i", type=int, default=100, dest="iters")
ap.add_argument("-nnz-per-row", type=int, default=11, dest="nnz_per_row")
ap.add_argument("-op", choices=["spmv", "spmm"], default="spmv")
ap.add_argument("-k", type=int, default=32)
ap.add_argument("--package", default="b200")
args = ap.parse_args()

name, timer, xp, sparse, _, on_device = select_package()
offsets = [d - args.nnz_per_row // 2 for d in range(args.nnz_per_row)]
A = sparse.diags([1] * args.nnz_per_row, offsets, shape=(args.n, args.n), format="csr", dtype=np.float64)
cols = 1 if args.op == "spmv" else args.k
x = xp.ones(args.n if args.op == "spmv" else (args.n, args.k))
y = xp.zeros(args.n if args.op == "spmv" else (args.n, args.k))


def product():
    global y
    if on_device:
        A.dot(x, out=y)
    else:
        y = A.dot(x)


product()  # warm-up (plan creation / page faults)
timer.start()
for _ in range(args.iters):
    product()
ms = timer.stop()
nnz = A.nnz
print(f"Iterations / sec: {args.iters / (ms / 1e3):.3f}")
print(f"[{name}] op={args.op} n={args.n} nnz={nnz} cols={cols} "
      f"{2 * nnz * cols * args.iters / (ms * 1e-3) / 1e9:.1f} GFLOP/s")
