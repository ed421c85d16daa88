"""SpMM Y = A @ X timing (dot_microbenchmark.py -op spmm): python tools/bench_spmm.py [n] [k ...]
A = banded 11 nnz/row and the 5-point Laplacian, fp64 and fp32; CUDA-event median over 20 launches, inputs
larger than L2.  Prints GFLOP/s and the algorithmic-bytes bandwidth (nnz*(sv+4) + 4(n+1) + 2*n*k*sv)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legate.sparse_b200 import _lib, _ops, gallery  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
ks = [int(v) for v in sys.argv[2:]] or [8, 32, 128]
side = int(round(n ** 0.5))
rows = []
for name, make in (("banded11", lambda dt: gallery.banded(n, 11, dtype=dt)),
                   ("lap5", lambda dt: gallery.laplacian_5pt(side, side, dtype=dt))):
    for dt in (np.float64, np.float32):
        A = make(dt)
        m = A.shape[0]
        for k in ks:
            X = torch.rand((A.shape[1], k), dtype=A._data.dtype, device="cuda")
            Y = torch.empty((m, k), dtype=A._data.dtype, device="cuda")
            times = {}
            for kern, kname in ((1, "row"), (2, "tile"), (3, "window"), (4, "tma")):
                _lib.lib.b2s_spmm_set_kernel(kern)
                for _ in range(3):
                    _ops.spmm(A._indptr, A._indices, A._data, X, Y, A.shape)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
                for s, e in ev:
                    s.record()
                    _ops.spmm(A._indptr, A._indices, A._data, X, Y, A.shape)
                    e.record()
                torch.cuda.synchronize()
                times[kname] = float(np.median([s.elapsed_time(e) for s, e in ev])) * 1e-3
            _lib.lib.b2s_spmm_set_kernel(0)
            t = min(times.values())   # the default dispatch picks by value type; see spmm.cu
            # spot check against the SpMV kernel on one column
            j = k // 2
            y = A @ X[:, j].contiguous()
            err = float((Y[:, j] - y).abs().max() / (y.abs().max() + 1e-30))
            sv = A.dtype.itemsize
            byts = A.nnz * (sv + 4) + 4 * (m + 1) + 2 * m * k * sv
            rows.append(dict(matrix=name, dtype=str(np.dtype(dt)), n=m, nnz=A.nnz, k=k, us=round(t * 1e6, 1), row_kernel_us=round(times['row'] * 1e6, 1), tile_kernel_us=round(times['tile'] * 1e6, 1), window_kernel_us=round(times['window'] * 1e6, 1), tma_window_kernel_us=round(times['tma'] * 1e6, 1),
                             gflops=round(2 * A.nnz * k / t / 1e9, 1), alg_gbs=round(byts / t / 1e9, 1),
                             col_err=err))
            print("SPMM", json.dumps(rows[-1]), flush=True)
            del X, Y
        del A
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/spmm_bench.json", "w"), indent=1)
