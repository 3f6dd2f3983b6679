"""Dense design products on the GPU box: correctness vs numpy on the full matrix and CUDA-event time per launch.
usage: python scripts/dense_bench.py [n] [p]      (default BASELINE config 2: 50000 x 5000, fp64, 2 GB)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuDenseDesignMatrix
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000
rng = np.random.default_rng(0)
X = rng.standard_normal((n, p))
ctx = _lib.Context.default()
for stream in (1, 0):
    ctx.set_option('dense_stream', stream)
    t0 = time.time()
    D = GpuDenseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx)
    up = time.time() - t0
    v, w = rng.standard_normal(p + 1), rng.standard_normal(n)
    c = X.mean(0)
    ref_dot = v[0] + X @ v[1:] - c @ v[1:]
    ref_t = np.concatenate(([w.sum()], X.T @ w - w.sum() * c))
    a, b = D.dot(v), D.Tdot(w)
    print('dense_stream=%d upload %.1fs  rel err dot %.2e Tdot %.2e' % (stream, up, np.linalg.norm(a - ref_dot) / np.linalg.norm(ref_dot),
                                                                       np.linalg.norm(b - ref_t) / np.linalg.norm(ref_t)), flush=True)
    gb = 8.0 * n * p / 1e9
    for what, passes in (('dot', 1), ('tdot', 1), ('op', 2)) + ((('fused_op', 1),) if stream else ()):
        for flush in (True, False):
            ms = D.time_kernel(what, reps=10, flush_l2=flush)
            print('   %-9s%s %8.1f us   X traffic %.2f GB -> %6.0f GB/s   (one-pass algorithmic 8np: %6.0f GB/s)'
                  % (what, ' ' if flush else '*', ms * 1e3, gb * passes, gb * passes / ms * 1e3, gb / ms * 1e3), flush=True)
    del D
ctx.set_option('dense_stream', 1)
