"""Index/value type choices (reference sparse/types.py:20-21 uses coord_ty=int64, nnz_ty=uint64)."""
import numpy as np

coord_ty = np.dtype(np.int32)   # column indices: 4 B instead of the reference's 8 B
nnz_ty = np.dtype(np.int64)     # per-row counts / SpGEMM indptr
float32 = np.dtype(np.float32)
float64 = np.dtype(np.float64)
