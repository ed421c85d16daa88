"""e2e host-vector SpMV (bench.py's `e2e`): y stored by the tiles straight into pinned host memory vs the copy-engine
D2H per stage, for several stage counts.  python tools/e2e_direct.py"""
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
ref = (A @ xh.cuda()).cpu().numpy()


def run(label):
    for _ in range(3):
        A.dot(x, out=y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        A.dot(x, out=y)
        ts.append(time.perf_counter() - t0)
    ok = np.array_equal(y, ref)
    print(f"{label:40s} median {np.median(ts)*1e3:.3f} ms  min {min(ts)*1e3:.3f} ms  -> {2*A.nnz/np.median(ts)/1e9:.1f} GFLOP/s  exact={ok}",
          flush=True)


for direct in ("1", "0"):
    os.environ["B2S_PIPE_DIRECT"] = direct
    for stages in (None, "16", "8", "6", "4", "2", "1"):
        if stages is None:
            os.environ.pop("B2S_PIPE_CHUNKS", None)
        else:
            os.environ["B2S_PIPE_CHUNKS"] = stages
        run(f"direct={direct} stages={stages or 'default'}")
os.environ.pop("B2S_PIPE_CHUNKS", None)
os.environ["B2S_PIPE_DIRECT"] = "1"
os.environ["B2S_PIPE_TRACE"] = "1"
A.dot(x, out=y)
