"""A/B of the bank-aware nnz ordering (option bank_permute) on the bench workload. Run on the GPU box.
usage: python scripts/bank_ab.py [C4|C4shard8|C3]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix

wl = sys.argv[1] if len(sys.argv) > 1 else 'C4'
n, p, dens = {'C4': (1000000, 100000, 0.001), 'C3': (100000, 10000, 0.01)}.get(wl.replace('shard8', ''), (1000000, 100000, 0.001))
blocks = range(bench.N_BLOCKS) if 'shard8' not in wl else range(bench.N_BLOCKS // 8)
t0 = time.time()
X, _ = bench.generate_rows(blocks, n, p, dens)
print('matrix', X.shape, X.nnz, 'generated in %.1fs' % (time.time() - t0), flush=True)
ctx = _lib.Context.default()
rng = np.random.default_rng(0)
w = rng.standard_normal(X.shape[0])
ref = {}
for binary in (True, False):
    for perm in (1, 0):
        ctx.set_option('bank_permute', perm)
        t0 = time.time()
        D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx, pattern_only=binary)
        up = time.time() - t0
        v = np.random.default_rng(1).standard_normal(D.shape[1])
        a, b = D.dot(v), D.Tdot(w)
        if perm == 1:
            ref[binary] = (a, b)
        else:
            ea = np.abs(a - ref[binary][0]).max() / np.abs(a).max()
            eb = np.abs(b - ref[binary][1]).max() / np.abs(b).max()
            print('  permuted vs canonical order: dot relerr %.2e, Tdot relerr %.2e' % (ea, eb), flush=True)
        bpn = 4 if binary else 12
        out = []
        for what in ('spmv_dot', 'spmv_tdot', 'op'):
            ms = D.time_kernel(what, reps=20, flush_l2=True)
            byt = bpn * X.nnz * (2 if what == 'op' else 1)
            out.append('%s %.1f us (%.0f GB/s)' % (what, ms * 1e3, byt / ms / 1e6))
        print('binary=%d bank_permute=%d upload %.1fs: %s' % (binary, perm, up, ', '.join(out)), flush=True)
        del D
ctx.set_option('bank_permute', 1)
