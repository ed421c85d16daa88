#!/usr/bin/env python
"""bench.py -- contract benchmark of the legate.sparse_b200 hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): CSR SpMV GFLOP/s and fraction of the HBM roofline.  Workload at N=1:
BASELINE.json configs[1] -- the 5-point Laplacian of examples/pde.py on a 3162 x 3162 interior grid
(N = 9,998,244 rows, nnz = 49,978,572), fp64 values, int32 indices, one SpMV per step.  At N>1 the grid
grows along y (3162 x 3162*N; weak scaling, 1-D row blocks) and every step includes the x halo exchange.

One JSON line on stdout (rank 0).  `value` = device-resident SpMV throughput (CUDA events, max over ranks);
`e2e` = the same metric through the public API with host (pinned) x and y, H2D + D2H inside the timed region;
`roofline` = algorithmic bytes / measured kernel time vs MEASURED_PEAKS.json; `cpu_baseline` = the CPU
oracle (OpenMP restatement of the reference's spmv_omp.cc) on this box's host cores.
`--impl reference` times that CPU implementation alone (the reference itself needs legate.core/Legion and
cannot be built here).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N1 = 3162  # interior grid edge: (nx-2) with nx = 3164 (L5)
METRIC = "csr_spmv_gflops"
WORKLOAD = "5-pt Laplacian (examples/pde.py operator) 3162x3162 interior grid per GPU, fp64 CSR SpMV, int32 indices"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes(nrows, ncols, nnz, sv=8, si=4, sp=4):
    """B = nnz*(sv+si) + (nrows+1)*sp + ncols*sv + nrows*sv."""
    return nnz * (sv + si) + (nrows + 1) * sp + ncols * sv + nrows * sv


# ----------------------------------------------------------------------------------------------------------
# host-side Laplacian assembly (numpy) for the CPU arms -- same operator as gallery.laplacian_5pt
# ----------------------------------------------------------------------------------------------------------
def laplacian_host(n1, n2):
    N = n1 * n2
    i = np.arange(N, dtype=np.int64)
    a = float((n1 + 1) ** 2)
    g = float((n2 + 1) ** 2)
    c = -2 * a - 2 * g
    cols = np.stack([i - n1, i - 1, i, i + 1, i + n1], axis=1)
    vals = np.tile(np.array([g, a, c, a, g]), (N, 1))
    valid = np.ones((N, 5), dtype=bool)
    valid[:, 0] = i >= n1
    valid[:, 1] = (i % n1) != 0
    valid[:, 3] = (i % n1) != n1 - 1
    valid[:, 4] = i < N - n1
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(valid.sum(axis=1), out=indptr[1:])
    return indptr, cols[valid].astype(np.int32), vals[valid], N


def calibrate_threads(orc, indptr, indices, data, x, y):
    """Pick the OpenMP thread count that runs the CPU SpMV fastest on this host (all logical CPUs is not always
    best: SMT siblings / cgroup quotas).  Returns the chosen count; the oracle is left configured with it."""
    ncpu = os.cpu_count() or 1
    cands = sorted({max(1, ncpu), max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 8)}, reverse=True)
    best, best_t = cands[0], float("inf")
    for n in cands:
        orc.set_num_threads(n)
        orc.spmv(indptr, indices, data, x, omp=True, out=y)
        t = time.perf_counter()
        orc.spmv(indptr, indices, data, x, omp=True, out=y)
        dt = time.perf_counter() - t
        if dt < best_t:
            best, best_t = n, dt
    orc.set_num_threads(best)
    return best


def cpu_spmv_rate(budget_s, indptr=None, indices=None, data=None, min_reps=3, max_reps=400):
    """OpenMP CPU oracle (reference spmv_omp.cc:36-45 restated) on the L5 matrix; returns GFLOP/s etc."""
    from oracle import oracle as orc

    orc.build()
    if indptr is None:
        indptr, indices, data, _ = laplacian_host(N1, N1)
    n = indptr.shape[0] - 1
    x = np.random.default_rng(0).random(n)
    y = np.zeros(n)
    orc.spmv(indptr, indices, data, x, omp=True, out=y)  # warm-up (page faults, thread pool)
    calibrate_threads(orc, indptr, indices, data, x, y)
    reps, t0 = 0, time.perf_counter()
    times = []
    while reps < max_reps and (reps < min_reps or time.perf_counter() - t0 < budget_s):
        t = time.perf_counter()
        orc.spmv(indptr, indices, data, x, omp=True, out=y)
        times.append(time.perf_counter() - t)
        reps += 1
    nnz = int(indptr[-1])
    med = float(np.median(times))
    return {"gflops": 2.0 * nnz / med / 1e9, "ms": med * 1e3, "reps": reps, "threads": orc.num_threads(),
            "nnz": nnz, "rows": n}


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (NVML; nvidia-smi fallback)
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
               0x10: "sync_boost"}

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# reference arm: CPU implementation of the path on this box's host cores
# ----------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        from oracle import oracle as orc

        orc.build()
    except Exception as exc:
        print(json.dumps({"impl": "reference", "unavailable": f"CPU oracle could not be built: {exc}"}), flush=True)
        return 0
    indptr, indices, data, n = laplacian_host(N1, N1)  # one shard of the workload = the bounded sample
    x = np.random.default_rng(0).random(n)
    nnz = int(indptr[-1])
    y = np.zeros(n)
    orc.spmv(indptr, indices, data, x, omp=True, out=y)
    calibrate_threads(orc, indptr, indices, data, x, y)
    tw = time.perf_counter()
    for _ in range(max(args.warmup, 1)):
        orc.spmv(indptr, indices, data, x, omp=True, out=y)
    per = (time.perf_counter() - tw) / max(args.warmup, 1)
    # bounded sample: keep the timed region under ~2 minutes of CPU work; report the steps actually run
    args.steps = int(max(1, min(args.steps, 120.0 / max(per, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.spmv(indptr, indices, data, x, omp=True, out=y)
    dt = time.perf_counter() - t0
    gf = 2.0 * nnz * args.steps / dt / 1e9
    threads = orc.num_threads()
    sample = f"{args.steps} SpMVs of one 3162x3162-grid shard ({n} rows, {nnz} nnz), OpenMP dynamic,128"
    line = {
        "impl": "reference", "metric": METRIC, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": n, "nnz": nnz,
                   "note": "reference's CPU leaf task (spmv_omp.cc) restated in oracle/oracle.c; the reference "
                           "itself needs legate.core/Legion and cannot be built here"},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch

    import legate.sparse_b200 as sparse
    from legate.sparse_b200 import _ops, gallery
    from legate.sparse_b200 import dist as bd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; legate.sparse_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bd.init_process_group("nccl")
    import torch.distributed as dist

    n1, n2 = N1, N1 * world
    Nglob = n1 * n2
    plan = bd.RowBlockPlan(Nglob, world)
    lo, hi = plan.rows(rank)
    local = gallery.laplacian_5pt(n1, n2, np.float64, row_lo=lo, row_hi=hi)
    nnz_local = local.nnz
    A = bd.dist_csr_array(local, (Nglob, Nglob), rank=rank, nranks=world)
    x_full = A.new_full_vector(np.float64)
    xl = A.local_view(x_full)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    xl.copy_(torch.rand(xl.shape[0], dtype=torch.float64, device="cuda", generator=g))
    y = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
    Al = A.local
    spmv_plan = Al._get_plan()
    xin = x_full[: Al.shape[1]]

    graphed = world > 1 and os.environ.get("B2S_BENCH_GRAPH", "1") != "0"

    def step():
        # halo exchange (overlapped with the interior tiles at N>1) + SpMV; at N>1 the step is replayed from a
        # CUDA graph because the NCCL send/recv call overhead on the host (~150 us) exceeds the 141 us kernel
        if graphed:
            A.dot_graphed(x_full, y)
        else:
            A.dot(x_full, out=y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    t_start.record()
    for s, e in kern_ev:
        s.record()
        step()
        e.record()
    t_end.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = t_start.elapsed_time(t_end)
    kern_ms = float(np.mean([s.elapsed_time(e) for s, e in kern_ev]))
    if world > 1:
        # at N>1 the per-step events bracket exchange + tiles; time the bare kernel separately for the roofline
        torch.cuda.synchronize()
        ke = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for s, e in ke:
            s.record()
            _ops.spmv(Al.indptr, Al.indices, Al.data, xin, y, Al.shape, plan=spmv_plan)
            e.record()
        torch.cuda.synchronize()
        kern_ms = float(np.mean([s.elapsed_time(e) for s, e in ke]))
    stats = torch.tensor([elapsed_ms, kern_ms], dtype=torch.float64, device="cuda")
    nnz_t = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(nnz_t, op=dist.ReduceOp.SUM)
    elapsed_ms, kern_ms = float(stats[0]), float(stats[1])
    nnz_glob = int(nnz_t[0])
    ms_per_step = elapsed_ms / args.steps
    value = 2.0 * nnz_glob / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: public API, host (pinned) vectors in and out, every step -----------------------------------
    x_host = torch.empty(Al.shape[1] if world == 1 else xl.shape[0], dtype=torch.float64).pin_memory()
    x_host.copy_(xl.cpu() if world > 1 else xin.cpu())
    y_host = torch.empty(hi - lo, dtype=torch.float64).pin_memory()
    xh_np, yh_np = x_host.numpy(), y_host.numpy()

    def e2e_step():
        if world == 1:
            Al.dot(xh_np, out=yh_np)  # H2D x, kernel, D2H y -- the call a user makes
        else:
            xl.copy_(x_host, non_blocking=True)
            step()
            y_host.copy_(y, non_blocking=False)

    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_gf = 2.0 * nnz_glob * e2e_steps / float(e2e_t[0]) / 1e9
    assert np.isfinite(yh_np[:1000]).all()

    peak, peak_src = measured_peak()
    B = alg_bytes(hi - lo, Al.shape[1] if world == 1 else (hi - lo) + A.recv_elems, nnz_local)
    achieved = B / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank != 0:
        return 0

    # ---- CPU baseline on rank 0 (N=1 only; bounded sample) ------------------------------------------------
    cpu = None
    extra = {}
    if world == 1 and not args.no_cpu:
        ip, ix, dv = (t.cpu().numpy() for t in (Al.indptr, Al.indices, Al.data))
        try:
            r = cpu_spmv_rate(args.cpu_budget, ip, ix, dv)
            cpu = {"value": r["gflops"], "unit": "GFLOP/s", "cores": r["threads"], "kind": "port",
                   "sample": f"{r['reps']} full SpMVs of the same matrix ({r['rows']} rows, {r['nnz']} nnz), "
                             f"median {r['ms']:.1f} ms, OpenMP oracle (reference spmv_omp.cc restated)"}
        except Exception as exc:  # the GPU numbers above must survive a broken host toolchain
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "port", "sample": f"unavailable: {exc}"}
        try:  # scipy (the reference tests' oracle): single-threaded csr_matvec, for context
            import scipy.sparse as sp

            S = sp.csr_array((dv, ix, ip), shape=Al.shape)
            xs = xin.cpu().numpy()
            S @ xs
            ts = []
            for _ in range(3):
                t = time.perf_counter(); S @ xs; ts.append(time.perf_counter() - t)
            extra["scipy_gflops_1thread"] = 2.0 * nnz_local / min(ts) / 1e9
            extra["host_cpu_count"] = os.cpu_count()
        except Exception as exc:  # pragma: no cover
            extra["scipy_error"] = str(exc)

    line = {
        "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_rows": Nglob, "global_nnz": nnz_glob, "rows_per_gpu": hi - lo,
                   "index_bytes": 4, "indptr_bytes": 4, "partition": f"1-D row blocks x{world}",
                   "x_exchange": A.exchange_mode, "halo_elems_per_rank": A.recv_elems,
                   "exchange_overlapped_with_interior_tiles": bool(world > 1 and os.environ.get("B2S_OVERLAP", "0") == "1"),
                   "step_replayed_from_cuda_graph": bool(graphed),
                   "l2": "inputs larger than L2 (matrix stream 600 MB + x/y 160 MB per step vs 126 MB L2); no flush",
                   "tile_config": int(spmv_plan.config), "kernel_family": "rowgroup" if spmv_plan.rowgroup else "tma-tiles",
                   "x_lines_per_warp_gather": spmv_plan.lines_per_warp},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": B,
                     "kernel": "b2s::spmv_tma_kernel<double,int,int,4,4,2,6>", "kernel_ms": kern_ms},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_gf, "unit": "GFLOP/s", "h2d_bytes_per_step": int(x_host.numel() * 8) * world,
                "d2h_bytes_per_step": int(y_host.numel() * 8) * world, "steps": e2e_steps,
                "api": "csr_array.dot(x_host, out=y_host)" if world == 1 else "shard copy-in + dist_csr_array.dot + copy-out"},
        "clocks": clocks,
        "gpu_launches": args.steps * world,
    }
    if world == 1 and not args.no_extras:
        line["extras"] = other_rows_of_the_path(torch, gallery, peak)
    line.update(extra)
    print(json.dumps(line), flush=True)
    return 0


def other_rows_of_the_path(torch, gallery, peak):
    """The other hot-path rows of SURVEY 8 at N=1, so one bench line records them all (bounded: ~6 s).
    CG: examples/pde.py -nx 4096 -ny 4096 -throughput -max_iter 300 (BASELINE config 3).
    SpGEMM: examples/spgemm_microbenchmark.py shape (banded, 11 nnz/row) at n = 1M."""
    from legate.sparse_b200 import linalg

    out = {}
    try:
        A = gallery.laplacian_5pt(4094, 4094, np.float64)
        b = torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
        linalg.cg(A, b, tol=1e-10, maxiter=30)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        x, iters = linalg.cg(A, b, tol=1e-10, maxiter=300)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        N, nnz = A.shape[0], A.nnz
        fused_bytes = nnz * 12 + N * 20 + 9 * 8 * N
        out["cg_pde4096"] = {"iters": iters, "it_per_s": iters / (ms * 1e-3), "us_per_iter": ms / iters * 1e3,
                             "model_bytes_per_iter": fused_bytes,
                             "frac_of_hbm_peak": fused_bytes * iters / (ms * 1e-3) / 1e9 / peak}
        del A, b, x
    except Exception as exc:  # pragma: no cover
        out["cg_error"] = str(exc)
    try:
        B = gallery.banded(1_000_000, 11, np.float64)
        C = B @ B
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            C = B @ B
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        info = C.spgemm_info
        out["spgemm_banded1m"] = {"ms": min(ts) * 1e3, "products": info["products"], "nnz_c": info["nnz"],
                                  "gflops": 2 * info["products"] / min(ts) / 1e9}
    except Exception as exc:  # pragma: no cover
        out["spgemm_error"] = str(exc)
    try:
        # SpMM: examples/dot_microbenchmark.py -op spmm -k 32 shape at n = 4M, fp64
        n, k = 4_000_000, 32
        B = gallery.banded(n, 11, np.float64)
        X = torch.rand((n, k), dtype=torch.float64, device="cuda")
        Y = torch.empty((n, k), dtype=torch.float64, device="cuda")
        for _ in range(3):
            B.dot(X, out=Y)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for s, e in ev:
            s.record()
            B.dot(X, out=Y)
            e.record()
        torch.cuda.synchronize()
        t = float(np.median([s.elapsed_time(e) for s, e in ev])) * 1e-3
        byts = B.nnz * 12 + 4 * (n + 1) + 2 * n * k * 8
        out["spmm_banded4m_k32"] = {"us": t * 1e6, "gflops": 2 * B.nnz * k / t / 1e9, "algorithmic_bytes": byts,
                                    "frac_of_hbm_peak": byts / t / 1e9 / peak}
        del B, X, Y
    except Exception as exc:  # pragma: no cover
        out["spmm_error"] = str(exc)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline sampling (rank 0, N=1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the CG / SpGEMM side measurements at N=1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
