"""Shared helpers of the example drivers: package switch and timers.

Plays the role of the reference's examples/benchmark.py (`--package legate|cupy|scipy`, one timer class per
package, :18-151): here the choices are `b200` (this repository, CUDA-event timer) and `scipy` (host,
perf_counter timer), so every example can be run against the CPU baseline with the same script.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class HostTimer:
    def start(self):
        self._t = time.perf_counter()

    def stop(self):
        """milliseconds since start()"""
        return (time.perf_counter() - self._t) * 1e3


class DeviceTimer:
    def __init__(self):
        import torch

        self._torch = torch

    def start(self):
        self._a = self._torch.cuda.Event(enable_timing=True)
        self._b = self._torch.cuda.Event(enable_timing=True)
        self._a.record()

    def stop(self):
        self._b.record()
        self._b.synchronize()
        return self._a.elapsed_time(self._b)


def select_package(argv=None):
    """Returns (name, timer, np_like, sparse, linalg, on_device). `np_like.ones/zeros` build dense vectors in
    the memory the chosen package computes from (device tensors for b200, numpy for scipy)."""
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--package", default="b200", choices=["b200", "scipy"])
    args, _ = ap.parse_known_args(argv)
    if args.package == "b200":
        import torch

        import legate.sparse_b200 as sparse
        from legate.sparse_b200 import linalg

        class DeviceArrays:
            float64 = "float64"

            @staticmethod
            def ones(n):
                return torch.ones(n, dtype=torch.float64, device="cuda")

            @staticmethod
            def zeros(n):
                return torch.zeros(n, dtype=torch.float64, device="cuda")

        return "b200", DeviceTimer(), DeviceArrays, sparse, linalg, True
    import numpy as np
    import scipy.sparse as sparse
    from scipy.sparse import linalg

    return "scipy", HostTimer(), np, sparse, linalg, False
