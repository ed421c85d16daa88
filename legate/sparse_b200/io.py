"""`mmread`: MatrixMarket coordinate files -> `coo_array`.

Host-side restatement of the reference's single-CPU-task reader (sparse/io.py:24-51 ->
READ_MTX_TO_COO, src/sparse/io/mtx_to_coo.cc:47-137): `real|pattern|integer` x `general|symmetric`,
1-based -> 0-based, off-diagonal entries of symmetric files mirrored right after the original,
values always float64 and coordinates int64 (mtx_to_coo.cc:28-29).
"""
from __future__ import annotations

import numpy as np

from .coo import coo_array


def mmread(source):
    with open(source, "r") as f:
        head = f.readline().split()
        if len(head) < 5 or head[0] != "%%MatrixMarket":
            raise ValueError("Unknown header of MatrixMarket")
        if head[1] != "matrix" or head[2] != "coordinate":
            raise ValueError("must be a coordinate matrix")
        field, symmetry = head[3], head[4]
        if field not in ("real", "pattern", "integer"):
            raise ValueError(f"unknown field {field}")
        if symmetry not in ("general", "symmetric"):
            raise ValueError(f"unknown symmetry {symmetry}")
        line = f.readline()
        while line and line.lstrip().startswith("%"):
            line = f.readline()
        dims = line.split()
        m, n, lines = int(dims[0]), int(dims[1]), int(dims[2])
        body = f.read().split()
    per = 2 if field == "pattern" else 3
    toks = np.array(body[: lines * per]).reshape(lines, per) if lines else np.empty((0, per), dtype=str)
    ci = toks[:, 0].astype(np.int64)
    cj = toks[:, 1].astype(np.int64)
    if field == "pattern":
        v = np.ones(lines, dtype=np.float64)
    elif field == "integer":
        v = toks[:, 2].astype(np.int64).astype(np.float64)
    else:
        v = toks[:, 2].astype(np.float64)
    if symmetry == "symmetric":
        off = ci != cj
        # mirrored entry sits right after its original (mtx_to_coo.cc:119-125)
        reps = np.where(off, 2, 1)
        idx = np.repeat(np.arange(lines), reps)
        second = np.zeros(idx.shape[0], dtype=bool)
        second[1:] = idx[1:] == idx[:-1]
        rows = np.where(second, cj[idx], ci[idx]) - 1
        cols = np.where(second, ci[idx], cj[idx]) - 1
        vals = v[idx]
    else:
        rows, cols, vals = ci - 1, cj - 1, v
    return coo_array((vals, (rows, cols)), shape=(m, n))
