import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from legate.sparse_b200 import gallery, linalg, _ops, runtime
A = gallery.laplacian_5pt(1022, 1022, np.float64)
N = A.shape[0]
b = torch.ones(N, dtype=torch.float64, device="cuda")
_ = A.dot(torch.zeros(N, dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); x, it = linalg.cg(A, b, tol=1e-10, maxiter=50); torch.cuda.synchronize()
    print(f"cg call {rep}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
os.environ["B2S_CG_GRAPH"] = "0"
for rep in range(2):
    t0 = time.perf_counter(); x, it = linalg.cg(A, b, tol=1e-10, maxiter=50); torch.cuda.synchronize()
    print(f"cg nograph call {rep}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
# isolate pieces
t0=time.perf_counter(); s=torch.cuda.Stream(); torch.cuda.synchronize(); print("stream", 1e3*(time.perf_counter()-t0))
t0=time.perf_counter(); g=torch.cuda.CUDAGraph(); 
with torch.cuda.graph(g, stream=s):
    b.add_(1.0)
torch.cuda.synchronize(); print("capture trivial", 1e3*(time.perf_counter()-t0))
