"""Small driver for ncu: dense-design kernels on BASELINE-config-2 / config-5 shaped matrices.
usage: python scripts/prof_dense.py stream|fisher|batch"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuDenseDesignMatrix
what = sys.argv[1] if len(sys.argv) > 1 else 'stream'
ctx = _lib.Context.default()
rng = np.random.default_rng(0)
if what == 'batch':
    n, p = 200_000, 2_000
else:
    n, p = 50_000, 5_000
X = rng.standard_normal((n, p))
D = GpuDenseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx)
if what == 'stream':
    for k in ('dot', 'tdot', 'fused_op'):
        print(k, D.time_kernel(k, reps=2, flush_l2=False) * 1e3, 'us', flush=True)
elif what == 'fisher':
    from bayesbridge_b200.reg_coef_sampler import generate_gaussian_with_weight
    _, st = generate_gaussian_with_weight(D, np.full(n, 0.5), np.ones(p + 1), np.zeros(p + 1), return_stats=True)
    print(st, flush=True)
else:
    _lib.check(_lib.load().bb_batch_init(D._mat, 16))
    _lib.check(_lib.load().bb_batch_set_obs_prec(D._mat, _lib.dptr(np.ascontiguousarray(rng.random((16, n)) * 0.25))))
    print('batch_op', D.time_kernel('batch_op', reps=2, flush_l2=False) * 1e3, 'us', flush=True)
