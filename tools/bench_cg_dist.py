"""Row-sharded CG throughput (torchrun, one rank per GPU): python tools/bench_cg_dist.py [nx] [max_iter] [weak]
strong: global grid (nx-2)^2 split over the ranks; weak: (nx-2) x (nx-2)*world."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import dist as bd, gallery  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 300
weak = len(sys.argv) > 3 and sys.argv[3] == "weak"
bd.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
n1 = nx - 2
n2 = n1 * world if weak else n1
N = n1 * n2
plan = bd.RowBlockPlan(N, world)
lo, hi = plan.rows(rank)
local = gallery.laplacian_5pt(n1, n2, np.float64, row_lo=lo, row_hi=hi)
A = bd.dist_csr_array(local, (N, N))
b = torch.ones(hi - lo, dtype=torch.float64, device="cuda")
bd.cg(A, b, tol=1e-10, maxiter=30)
dist.barrier(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
x, iters = bd.cg(A, b, tol=1e-10, maxiter=max_iter)
e.record()
dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t[0])
    print(f"CGDIST world={world} {'weak' if weak else 'strong'} nx={nx} N={N} exchange={A.exchange_mode}: {iters} iters "
          f"{ms:.2f} ms -> {iters/(ms*1e-3):.1f} it/s ({ms/iters*1e3:.1f} us/iter)")
dist.destroy_process_group()
