"""Tiny driver for ncu captures: python tools/prof_spmv.py <workload> <cfg> <waves> [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import _lib, _ops, gallery  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "l5"
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
waves = int(sys.argv[3]) if len(sys.argv) > 3 else 0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
if wl == "l5":
    A = gallery.laplacian_5pt(3162, 3162, np.float64)
elif wl == "banded":
    A = gallery.banded(10_000_000, 11, np.float64)
elif wl == "r32":
    A = gallery.random_fixed(10_000_000, 10_000_000, 32, np.float32)
elif wl == "b32f32":
    A = gallery.banded(10_000_000, 32, np.float32)
elif wl == "l5f32":
    A = gallery.laplacian_5pt(3162, 3162, np.float32)
else:
    raise SystemExit("unknown workload")
_lib.check(_lib.lib.b2s_spmv_set_config(cfg, waves))  # cfg < 0: automatic
plan = A._get_plan()
x = torch.rand(A.shape[1], dtype=A.data.dtype, device="cuda")
y = torch.empty(A.shape[0], dtype=A.data.dtype, device="cuda")
for _ in range(iters):
    _ops.spmv(A.indptr, A.indices, A.data, x, y, A.shape, plan=plan)
torch.cuda.synchronize()
print("done", wl, cfg, waves, float(y.sum()))
