"""Small driver for ncu: a few SpMV launches on a C3-shaped (or C4-shaped with 'big') binary matrix."""
import sys, os
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix
ctx = _lib.Context.default()
wl = {"big": "C4", "shard8": "C4shard8"}.get(sys.argv[1] if len(sys.argv) > 1 else "", "C3")
valued = len(sys.argv) > 2 and sys.argv[2] == 'valued'
n, p, dens = bench.WORKLOADS[wl]
X, y = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
print(wl, X.shape, X.nnz, flush=True)
D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx, pattern_only=not valued)
for what in ('spmv_dot', 'spmv_tdot', 'op'):
    print(what, D.time_kernel(what, reps=3, flush_l2=False) * 1e3, 'us', flush=True)
