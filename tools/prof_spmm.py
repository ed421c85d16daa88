"""One launch of each SpMM kernel (row, tile) per k on the banded 11/row fp64 matrix -- the target of
`ncu --set full -k regex:spmm_ python tools/prof_spmm.py [n] [k ...]`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import _lib, _ops, gallery  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
ks = [int(v) for v in sys.argv[2:]] or [32, 128]
A = gallery.banded(n, 11, dtype=np.float64)
for k in ks:
    X = torch.rand((n, k), dtype=torch.float64, device="cuda")
    Y = torch.empty((n, k), dtype=torch.float64, device="cuda")
    for kern in (1, 2, 3, 4):
        _lib.lib.b2s_spmm_set_kernel(kern)
        _ops.spmm(A._indptr, A._indices, A._data, X, Y, A.shape)
        torch.cuda.synchronize()
    del X, Y
