"""GPU sweep of the SpMV tile configurations on the BASELINE workloads (run under gpurun).

Prints one line per (workload, config, waves): time, GFLOP/s, achieved GB/s and fraction of the measured
HBM peak.  Development tool -- bench.py is the contract benchmark."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import legate.sparse_b200 as sparse  # noqa: E402
from legate.sparse_b200 import _lib, _ops, gallery  # noqa: E402

PEAK = 6583.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def alg_bytes(A):
    sv = A.dtype.itemsize
    si = A.indices.element_size()
    spt = A.indptr.element_size()
    return A.nnz * (sv + si) + (A.shape[0] + 1) * spt + A.shape[1] * sv + A.shape[0] * sv


def time_spmv(A, x, y, plan, iters=20, warm=3):
    for _ in range(warm):
        _ops.spmv(A.indptr, A.indices, A.data, x, y, A.shape, plan=plan)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        _ops.spmv(A.indptr, A.indices, A.data, x, y, A.shape, plan=plan)
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3


def run(name, A, cfgs, waves_list, out):
    x = torch.rand(A.shape[1], dtype=A.data.dtype, device="cuda")
    y = torch.empty(A.shape[0], dtype=A.data.dtype, device="cuda")
    B = alg_bytes(A)
    t_med, t_min = time_spmv(A, x, y, None)
    line = f"{name:28s} plan-free          med {t_med*1e6:8.1f} us  {2*A.nnz/t_med/1e9:8.1f} GF/s  {B/t_med/1e9:7.1f} GB/s  frac {B/t_med/1e9/PEAK:.3f}"
    print(line); out.write(line + "\n"); out.flush()
    yref = y.clone()
    for cfg in cfgs:
        for waves in waves_list:
            _lib.check(_lib.lib.b2s_spmv_set_config(cfg, waves))
            plan = _ops.spmv_plan(A.indptr, A.indices, A.shape, A.nnz, A.dtype)
            plan.set_kernel(False)
            flavors = [None]
            if plan.tma and os.environ.get("SWEEP_FLAVORS"):
                flavors = [None] + [int(f) for f in os.environ["SWEEP_FLAVORS"].split(",")]
            for fl in flavors:
                if fl is not None:
                    plan.set_flavor(fl)
                t_med, t_min = time_spmv(A, x, y, plan)
                scale = float(yref.abs().max()) + 1e-30
                ok = bool(((y - yref).abs().max() / scale) < (1e-5 if A.dtype == np.float32 else 1e-12))
                tag = "auto:" + ("short" if plan.short_rows else "uni" if plan.uniform else "gen") if fl is None else f"flavor{fl}"
                line = (f"{name:28s} cfg {cfg} waves {waves:2d} {tag:10s} med {t_med*1e6:8.1f} us  min {t_min*1e6:8.1f} us  "
                        f"{2*A.nnz/t_med/1e9:8.1f} GF/s  {B/t_med/1e9:7.1f} GB/s  frac {B/t_med/1e9/PEAK:.3f}  ok={ok}")
                print(line); out.write(line + "\n"); out.flush()
    _lib.check(_lib.lib.b2s_spmv_set_config(-1, 0))
    auto = A._get_plan()
    t_med, t_min = time_spmv(A, x, y, auto)
    line = (f"{name:28s} AUTO cfg {auto.config} rowgroup={auto.rowgroup} lines/warp={auto.lines_per_warp:.1f}  med {t_med*1e6:8.1f} us  "
            f"{2*A.nnz/t_med/1e9:8.1f} GF/s  {B/t_med/1e9:7.1f} GB/s  frac {B/t_med/1e9/PEAK:.3f}")
    print(line); out.write(line + "\n"); out.flush()


def copy_bw(out):
    n = 1 << 28
    a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        b.copy_(a)
    e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) * 1e-3 / 10
    line = f"torch copy 1 GiB x2: {2*n*4/t/1e9:.1f} GB/s (MEASURED_PEAKS {PEAK})"
    print(line); out.write(line + "\n")


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    which = sys.argv[1:] or ["l5", "banded", "r32", "r32w", "r32f64"]
    cfgs = list(range(int(_lib.lib.b2s_spmv_num_configs())))
    if os.environ.get("SWEEP_CFGS"):
        cfgs = [int(c) for c in os.environ["SWEEP_CFGS"].split(",")]
    waves = [0, 2, 4]
    with open("gpurun_out/sweep_spmv.txt", "a") as out:
        out.write(f"# {time.ctime()} {torch.cuda.get_device_name(0)} {sparse.runtime.device_info()}\n")
        copy_bw(out)
        if "l5" in which:
            run("L5 fp64 5pt 3162^2", gallery.laplacian_5pt(3162, 3162, np.float64), cfgs, waves, out)
        if "banded" in which:
            run("banded11 fp64 n=10M", gallery.banded(10_000_000, 11, np.float64), cfgs, [0], out)
        if "r32" in which:
            run("R32 fp32 random 10M", gallery.random_fixed(10_000_000, 10_000_000, 32, np.float32), cfgs, [0], out)
        if "r32w" in which:
            run("R32 fp32 window 64K", gallery.random_fixed(10_000_000, 10_000_000, 32, np.float32, window=65536), cfgs, [0], out)
        if "l5f32" in which:
            run("L5 fp32 5pt 3162^2", gallery.laplacian_5pt(3162, 3162, np.float32), cfgs, [0], out)
        if "b32f32" in which:
            run("banded32 fp32 n=10M", gallery.banded(10_000_000, 32, np.float32), cfgs, [0], out)
        if "r4" in which:     # scattered AND short rows (what a column block of a weak-scaled R32 shard looks like)
            run("R4 fp32 random 10M x 40M", gallery.random_fixed(10_000_000, 40_000_000, 4, np.float32), cfgs, [0], out)
            run("R4 fp32 random 10M x 10M", gallery.random_fixed(10_000_000, 10_000_000, 4, np.float32), cfgs, [0], out)
        if "r4f64" in which:
            run("R4 fp64 random 10M x 5M", gallery.random_fixed(10_000_000, 5_000_000, 4, np.float64), cfgs, [0], out)
        if "r32f64" in which:
            run("R32 fp64 random 10M", gallery.random_fixed(10_000_000, 10_000_000, 32, np.float64), cfgs, [0], out)
