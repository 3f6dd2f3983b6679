"""Kernel timings on C3-/C4-shaped matrices (prints). Run on the GPU box."""
import sys, os, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix

ctx = _lib.Context.default()
rng = np.random.default_rng(0)
sizes = [(100000, 20000, 0.005)]
if len(sys.argv) > 1 and sys.argv[1] == 'big':
    sizes.append((1000000, 100000, 0.001))
for (n, p, dens) in sizes:
    nnz = int(n * p * dens)
    rows = rng.integers(0, n, nnz)
    cols = (rng.beta(0.5, 20, nnz) * p).astype(np.int64) % p
    X = sp.csr_matrix((np.ones(nnz), (rows, cols)), shape=(n, p)); X.sum_duplicates(); X.data[:] = 1.0
    print("matrix", n, p, X.nnz, flush=True)
    for binary in (True, False):
        for stage in (1, 0, 2, 3):
            ctx.set_option('spmv_stage', stage)
            t0 = time.time()
            D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx, pattern_only=binary)
            t1 = time.time()
            nz = X.nnz; bpn = 4 if binary else 12
            for what in ('spmv_dot', 'spmv_tdot', 'dot', 'tdot', 'op'):
                ms = D.time_kernel(what, reps=20, flush_l2=True)
                ms2 = D.time_kernel(what, reps=20, flush_l2=False)
                byt = bpn * nz * (2 if what == 'op' else 1)
                print(f"n={n} nnz={nz} binary={binary} stage={stage} {what}: {ms*1e3:.1f} us cold ({byt/ms/1e6:.0f} GB/s), {ms2*1e3:.1f} us warm ({byt/ms2/1e6:.0f} GB/s)  [upload {t1-t0:.1f}s]", flush=True)
            del D
