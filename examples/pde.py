#!/usr/bin/env python
"""2-D Poisson problem solved with CG -- the workload of the reference's examples/pde.py.

Dirichlet problem on [0,1] x [-0.5,0.5] with the 5-point second-order stencil on an nx x ny grid (interior
(nx-2) x (ny-2) unknowns, column-major numbering), right-hand side
    b = sin(pi x) cos(pi y) + sin(5 pi x) cos(5 pi y)
whose exact solution is  -1/(2 pi^2) sin(pi x) cos(pi y) - 1/(50 pi^2) sin(5 pi x) cos(5 pi y).
`-throughput` solves with b = 1 for exactly `-max_iter` iterations and prints "Iterations / sec"
(the mode the reference's published logs use); otherwise the solve runs to tol = 1e-10 and the error
against the exact solution is printed.

    python examples/pde.py -nx 4096 -ny 4096 -throughput -max_iter 300 [--package scipy]
"""
import argparse
import sys

import numpy as np

from common import select_package

ap = argparse.ArgumentParser()
ap.add_argument("-nx", type=int, default=101)
ap.add_argument("-ny", type=int, default=101)
ap.add_argument("-throughput", action="store_true")
ap.add_argument("-max_iter", type=int, default=None)
ap.add_argument("--package", default="b200")
args = ap.parse_args()
if args.throughput and args.max_iter is None:
    sys.exit("-throughput needs -max_iter")

name, timer, xp, sparse, linalg, on_device = select_package()
nx, ny = args.nx, args.ny
mx, my = nx - 2, ny - 2            # interior points per direction
hx, hy = 1.0 / (nx - 1), 1.0 / (ny - 1)
xs = np.linspace(0.0, 1.0, nx)
ys = np.linspace(-0.5, 0.5, ny)


def rhs_and_exact():
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    f = np.sin(np.pi * X) * np.cos(np.pi * Y) + np.sin(5 * np.pi * X) * np.cos(5 * np.pi * Y)
    u = (-np.sin(np.pi * X) * np.cos(np.pi * Y) / (2 * np.pi**2)
         - np.sin(5 * np.pi * X) * np.cos(5 * np.pi * Y) / (50 * np.pi**2))
    return f, u


def laplacian():
    """d2/dx2 + d2/dy2 on the interior grid, unknowns numbered with x fastest.  Five diagonals; the +-1
    couplings vanish across grid lines (those entries are explicit zeros that the DIA->CSR conversion drops)."""
    cx, cy = 1.0 / hx**2, 1.0 / hy**2
    n = mx * my
    main = np.full(n, -2.0 * (cx + cy))
    near = np.full(n - 1, cx)
    near[mx - 1 :: mx] = 0.0
    far = np.full(n - mx, cy)
    M = sparse.diags([far, near, main, near, far], [-mx, -1, 0, 1, mx], dtype=np.float64)
    return M.tocsc().T if name == "b200" else M.tocsr()


A = laplacian()
if args.throughput:
    b = xp.ones(mx * my)
else:
    f, exact = rhs_and_exact()
    b = f[1:-1, 1:-1].flatten("F")

_ = A.dot(xp.zeros(A.shape[1]))  # warm-up: builds the SpMV plan before the clock starts
timer.start()
if args.throughput:
    if name == "b200":
        sol, iters = linalg.cg(A, b, tol=1e-10, maxiter=args.max_iter)
    else:
        sol, info = linalg.cg(A, b, rtol=0.0, atol=1e-10, maxiter=args.max_iter)
    ms = timer.stop()
    print(f"Iterations / sec: {args.max_iter / (ms / 1e3):.3f}")
    sys.exit(0)
if name == "b200":
    sol, iters = linalg.cg(A, b, tol=1e-10)
else:
    sol, info = linalg.cg(A, b, rtol=0.0, atol=1e-10)
ms = timer.stop()
sol = np.asarray(sol)
assert np.allclose(np.asarray(A @ sol), b)
print(f"Total time: {ms:.3f} ms")
u = np.zeros((nx, ny))
u[1:-1, 1:-1] = sol.reshape((mx, my), order="F")
print(f"Iterative method error: {np.sqrt(np.sum((u - exact) ** 2))}")
