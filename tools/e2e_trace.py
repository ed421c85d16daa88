"""Timeline of the host-vector SpMV pipeline (B2S_PIPE_TRACE=1 python tools/e2e_trace.py): when each chunk's H2D,
tiles and D2H ran, plus the whole-call time over 20 calls.  Explains the e2e number of bench.py."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import gallery  # noqa: E402

A = gallery.laplacian_5pt(3162, 3162, np.float64)
xh = torch.rand(A.shape[1], dtype=torch.float64).pin_memory()
yh = torch.empty(A.shape[0], dtype=torch.float64).pin_memory()
x, y = xh.numpy(), yh.numpy()
os.environ.pop("B2S_PIPE_TRACE", None)
for _ in range(3):
    A.dot(x, out=y)
torch.cuda.synchronize()
os.environ["B2S_PIPE_TRACE"] = "1"      # one traced call (the library reads the variable per call)
A.dot(x, out=y)
os.environ.pop("B2S_PIPE_TRACE", None)
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    A.dot(x, out=y)
    ts.append(time.perf_counter() - t0)
print(f"e2e call: median {np.median(ts)*1e3:.3f} ms  min {min(ts)*1e3:.3f} ms  -> {2*A.nnz/np.median(ts)/1e9:.1f} GFLOP/s")
