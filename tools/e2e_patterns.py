"""Host-vector SpMV (bench.py's `e2e`) under different stage patterns / copy alignments of the pipeline.
python tools/e2e_patterns.py   -> one line per variant (wall clock median of 30 products, exactness checked)."""
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


def run(label, reps=30):
    for _ in range(3):
        A.dot(x, out=y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        A.dot(x, out=y)
        ts.append(time.perf_counter() - t0)
    ok = np.array_equal(y, ref)
    print(f"{label:46s} median {np.median(ts)*1e3:.3f} ms  min {min(ts)*1e3:.3f} ms  -> {2*A.nnz/np.median(ts)/1e9:.1f} GFLOP/s  exact={ok}",
          flush=True)


PATTERNS = ["1,3,4,4,3,1", "2,2,2,2,2,2,2,2", "1,1,2,2,2,2,2,2,1,1", "1,2,2,2,2,2,2,2,1", "1,2,3,4,3,2,1", "1,1,2,4,4,2,1,1",
            "2,4,4,4,2", "1,2,4,4,4,1", "1,3,3,3,3,3", "1,2,2,3,3,3,2", "1,1,1,1,2,2,2,2,2,2", "4,4,4,4", "1,3,4,4,4", "2,3,3,3,3,2"]
for align in ("0", "512", "8192"):
    os.environ["B2S_PIPE_ALIGN"] = align
    for pat in PATTERNS if align == "0" else PATTERNS[:4]:
        os.environ["B2S_PIPE_PATTERN"] = pat
        run(f"align={align} pattern={pat}")
os.environ.pop("B2S_PIPE_ALIGN")
for pat in ("1,3,4,4,3,1", "2,2,2,2,2,2,2,2"):
    os.environ["B2S_PIPE_PATTERN"] = pat
    os.environ["B2S_PIPE_TRACE"] = "1"
    print("trace", pat, flush=True)
    A.dot(x, out=y)
    os.environ.pop("B2S_PIPE_TRACE")
