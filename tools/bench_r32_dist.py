"""BASELINE config 4: random CSR, 10M rows x 32 nnz/row per GPU, fp32, weak scaling; x all-gathered per SpMV.
torchrun: python tools/bench_r32_dist.py [rows_per_gpu] [steps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import dist as bd, gallery  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    bd.init_process_group("nccl")
rank = dist.get_rank() if world > 1 else 0
N = rows * world
local = gallery.random_fixed(rows, N, 32, np.float32, seed=1234 + rank)
A = bd.dist_csr_array(local, (N, N), rank=rank, nranks=world)
xf = A.new_full_vector(np.float32)
A.local_view(xf).copy_(torch.rand(rows, dtype=torch.float32, device="cuda"))
y = torch.empty(rows, dtype=torch.float32, device="cuda")
for _ in range(5):
    A.dot(xf, out=y)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(steps):
    A.dot(xf, out=y)
e.record()
torch.cuda.synchronize()
t = torch.tensor([s.elapsed_time(e) / steps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t[0])
if rank == 0:
    plan = A.local._get_plan()
    nnz = rows * 32 * world
    print(f"R32DIST world={world} rows/gpu={rows} exchange={A.exchange_mode} cfg={plan.config} scattered={plan.scattered}: "
          f"{ms*1e3:.1f} us/step -> {2*nnz/(ms*1e-3)/1e9:.1f} GFLOP/s total ({2*nnz/(ms*1e-3)/1e9/world:.1f} per GPU)")
if world > 1:
    dist.destroy_process_group()
