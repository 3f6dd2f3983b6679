"""A/B of the SpMV kernels on a bench workload (run on the GPU box): correctness against scipy on the full
matrix, upload time and CUDA-event time per launch for every variant.
usage: python scripts/spmv_ab.py [C4|C4shard8|C3|C1] [variants, e.g. 1,1nb,1np,0] [valued]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix

wl = sys.argv[1] if len(sys.argv) > 1 else 'C4'
variants = (sys.argv[2] if len(sys.argv) > 2 else '1,1nb,1np,0').split(',')
valued = len(sys.argv) > 3 and sys.argv[3] == 'valued'
n, p, dens = bench.WORKLOADS[wl]
t0 = time.time()
X, _ = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
print('matrix', wl, X.shape, X.nnz, 'generated in %.1fs' % (time.time() - t0), flush=True)
ctx = _lib.Context.default()
rng = np.random.default_rng(0)
w = rng.standard_normal(X.shape[0])
ref_dot = ref_tdot = v = None
for var in variants:
    ctx.set_option('spmv_variant', int(var[0]))
    ctx.set_option('spmv_bulk', 0 if 'nb' in var else 1)
    ctx.set_option('bank_permute', 0 if 'np' in var else (2 if 'p2' in var else 1))
    t0 = time.time()
    D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx, pattern_only=not valued)
    up = time.time() - t0
    if v is None:       # scipy reference on the host image (zero-variance columns are dropped at construction)
        Xm, c = D.X_main, D.column_offset
        v = rng.standard_normal(D.shape[1])
        t0 = time.time()
        ref_dot = v[0] + Xm @ v[1:] - c @ v[1:]
        ref_tdot = np.concatenate(([w.sum()], Xm.T @ w - w.sum() * c))
        print('scipy dot + Tdot: %.2fs' % (time.time() - t0), flush=True)
    a, b = D.dot(v), D.Tdot(w)
    ea = np.abs(a - ref_dot).max() / np.abs(ref_dot).max()
    eb = np.abs(b - ref_tdot).max() / np.abs(ref_tdot).max()
    a2, b2 = D.dot(v), D.Tdot(w)
    rep = np.array_equal(a, a2) and np.array_equal(b, b2)
    bpn = 12 if valued else 4
    out = []
    for what in ('spmv_dot', 'spmv_tdot', 'dot', 'tdot', 'op'):
        for flush in (True, False):
            ms = D.time_kernel(what, reps=20, flush_l2=flush)
            byt = bpn * X.nnz * (2 if what == 'op' else 1)
            out.append('%s%s %.1f us (%.0f GB/s)' % (what, '' if flush else '[warm]', ms * 1e3, byt / ms / 1e6))
    print('variant=%s upload %.1fs  max-rel-err vs scipy: dot %.2e Tdot %.2e  reproducible=%s\n   %s'
          % (var, up, ea, eb, rep, '\n   '.join(out)), flush=True)
    del D
ctx.set_option('spmv_variant', 1)
ctx.set_option('spmv_bulk', 1)
ctx.set_option('bank_permute', 1)
