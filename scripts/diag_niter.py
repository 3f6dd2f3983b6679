"""Diagnostic: CG residual history on the BASELINE-config-1 fixture, device vs oracle, under different chunking options."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cg_oracle as co
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix
from bayesbridge_b200.reg_coef_sampler import ConjugateGradientSampler
g = np.load('tests/golden/cg_c1_ref.npz')
X = sp.csr_matrix((np.ones(len(g['indices'])), g['indices'], g['indptr']), shape=tuple(g['shape']))
ctx = _lib.Context.default()
omega, pps, z, x0, sd = (g[k] for k in ('omega', 'pps', 'z', 'x0', 'sd'))
n, P = X.shape[0], X.shape[1] + 1
O = co.DesignOracle(X, True, True)
s = co.precond_scale_prior(pps, 1, sd)
np.random.seed(7); e1, e2 = np.random.randn(n), np.random.randn(P)
atol = 1e-5 * np.sqrt(P)
for opts in ({}, {'use_graph': 0}, {'cg_chunk': 1}, {'cg_fused': 0}, {'cg_fused': 0, 'use_graph': 0}, {'spmv_variant': 0}):
    for k, v in opts.items():
        ctx.set_option(k, v)
    D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx)
    out = []
    for K in (19, 20, 21, 22):
        ref, _ = co.cg_sample(O, omega, pps, z, x0, s, K, 0.0, e1, e2)
        coef, info = ConjugateGradientSampler(1).sample(D, omega, pps, z, x0.copy(), 'prior', sd, maxiter=K, atol=0.0, seed=7, return_stats=True)
        out.append('K=%d n_iter=%d rnorm/atol=%.4f err=%.1e' % (K, info['n_iter'], info['resid_norm'] / atol, np.linalg.norm(coef - ref) / np.linalg.norm(ref)))
    coef, info = ConjugateGradientSampler(1).sample(D, omega, pps, z, x0.copy(), 'prior', sd, maxiter=500, atol=atol, seed=7, return_stats=True)
    coef2, info2 = ConjugateGradientSampler(1).sample(D, omega, pps, z, x0.copy(), 'prior', sd, maxiter=500, atol=atol, seed=7, return_stats=True)
    print(opts, '| default rule: n_iter', info['n_iter'], 'again', info2['n_iter'], 'rnorm/atol %.4f' % (info['resid_norm'] / atol), '|', ' ; '.join(out), flush=True)
    for k in opts:
        ctx.set_option(k, {'use_graph': 1, 'cg_chunk': 0, 'cg_fused': 1, 'spmv_variant': 1}[k])
