"""Quick on-GPU sanity run (prints, does not assert): products, CSC bit-exactness, CG vs the reference
sampler with injected noise, PG / tilted-stable moments, a short chain, and kernel timings."""
import sys, os, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesbridge_b200 as bb
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix, GpuDenseDesignMatrix
from bayesbridge_b200.reg_coef_sampler import ConjugateGradientSampler

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
have_ref = os.path.isdir(REF)
if have_ref:
    sys.path.insert(0, REF)
    import bayesbridge as ref
    from bayesbridge.design_matrix import SparseDesignMatrix as RefSparse, DenseDesignMatrix as RefDense
    from bayesbridge.reg_coef_sampler.cg_sampler import ConjugateGradientSampler as RefCG

ctx = _lib.Context.default()
rng = np.random.default_rng(0)

def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

def rand_sparse(n, p, dens, binary, hot=True):
    X = sp.random(n, p, density=dens, format='csr', random_state=np.random.RandomState(1), dtype=np.float64)
    if hot:
        col = sp.csr_matrix((np.ones(n // 2), (rng.choice(n, n // 2, replace=False), np.full(n // 2, 3))), shape=(n, p))
        X = (X + col).tocsr()
    if binary:
        X.data[:] = 1.0
    return X

print("== products ==")
for (n, p, dens) in [(300, 40, 0.2), (20000, 700, 0.02), (60000, 3000, 0.004)]:
    for binary in (False, True):
        for slab in (0, 64, 1024):
            for stage in (1, 0):
                ctx.set_option('slab_width', slab); ctx.set_option('spmv_stage', stage)
                X = rand_sparse(n, p, dens, binary)
                for center in (False, True):
                    for icpt in (False, True):
                        D = GpuSparseDesignMatrix(X, center_predictor=center, add_intercept=icpt, ctx=ctx)
                        A = X.toarray() - (X.toarray().mean(0) if center else 0)
                        if icpt: A = np.hstack((np.ones((n, 1)), A))
                        v = rng.standard_normal(A.shape[1]); w = rng.standard_normal(n); wt = rng.random(n)
                        e1 = relerr(D.dot(v), A @ v); e2 = relerr(D.Tdot(w), A.T @ w)
                        e3 = relerr(D.compute_fisher_info(wt, diag_only=True), (A * A * wt[:, None]).sum(0))
                        bad = max(e1, e2, e3) > 1e-11
                        if bad or (center and icpt):
                            print(f"n={n} p={p} bin={binary} slab={slab} stage={stage} c={center} i={icpt}: dot {e1:.1e} tdot {e2:.1e} fisher {e3:.1e} {'BAD' if bad else ''}")
                        if center and icpt and slab == 0 and stage == 1:
                            ip, ix, dv = D.export_csc(); C = X.tocsc()
                            print("   csc bit-exact:", np.array_equal(ip, C.indptr), np.array_equal(ix, C.indices), np.array_equal(dv, C.data))
                        del D
ctx.set_option('slab_width', 0); ctx.set_option('spmv_stage', 1)

print("== dense products ==")
for (n, p) in [(200, 30), (5000, 1300)]:
    Xd = rng.standard_normal((n, p))
    for center in (False, True):
        for icpt in (False, True):
            D = GpuDenseDesignMatrix(Xd.copy(), center_predictor=center, add_intercept=icpt, ctx=ctx)
            A = Xd - (Xd.mean(0) if center else 0)
            if icpt: A = np.hstack((np.ones((n, 1)), A))
            v = rng.standard_normal(A.shape[1]); w = rng.standard_normal(n); wt = rng.random(n)
            print(f"dense n={n} p={p} c={center} i={icpt}: dot {relerr(D.dot(v), A@v):.1e} tdot {relerr(D.Tdot(w), A.T@w):.1e} fisher {relerr(D.compute_fisher_info(wt, True), (A*A*wt[:,None]).sum(0)):.1e}")

print("== CG vs reference sampler (injected noise) ==")
if have_ref:
    for (n, p, dens, dense) in [(5000, 400, 0.05, False), (20000, 2000, 0.01, False), (3000, 200, 1.0, True)]:
        if dense:
            X = rng.standard_normal((n, p)); Dg = GpuDenseDesignMatrix(X.copy(), center_predictor=True, add_intercept=True, ctx=ctx)
            Dr = RefDense(X.copy(), center_predictor=True, add_intercept=True)
        else:
            X = rand_sparse(n, p, dens, True); Dg = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx)
            Dr = RefSparse(X, use_mkl=False, center_predictor=True, add_intercept=True)
        P = p + 1
        omega = rng.random(n) * 0.25 + 0.01
        pps = np.concatenate(([0.0], 1 / (0.1 * rng.random(p) + 1e-3)))
        z = rng.standard_normal(P); x0 = 0.01 * rng.standard_normal(P); sd = np.ones(P) * 0.7
        for (maxiter, atol) in [(1, 0.0), (5, 0.0), (20, 0.0), (500, 1e-5 * np.sqrt(P)), (500, 1e-12 * np.sqrt(P))]:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                cr, ir = RefCG(1).sample(Dr, omega, pps, z, x0.copy(), 'prior', sd, maxiter=maxiter, atol=atol, seed=7)
                cg, ig = ConjugateGradientSampler(1).sample(Dg, omega, pps, z, x0.copy(), 'prior', sd, maxiter=maxiter, atol=atol, seed=7, return_stats=True)
            print(f"n={n} p={p} dense={dense} maxiter={maxiter} atol={atol:.1e}: rel {relerr(cg, cr):.2e} n_iter ref {ir['n_iter']} gpu {ig['n_iter']} conv {ir['converged']}/{ig['converged']} ms {ig['device_ms']:.2f}")
        del Dg

print("== PG ==")
from bayesbridge_b200.random import DevicePolyaGamma, DeviceTiltedStable
pg = DevicePolyaGamma(ctx, 5)
N = 400000
for b in (1, 2, 5):
    for c in (0.0, 0.01, 0.5, 2.0, 10.0, 50.0, 100.0):
        x = pg.rand_polyagamma(np.full(N, b, dtype=np.int32), np.full(N, c))
        m = b / 4 if c < 1e-5 else b / (2 * c) * np.tanh(c / 2)
        v = b / 24 if c < 1e-5 else b * (np.sinh(c) - c) / (4 * c ** 3 * np.cosh(c / 2) ** 2) if c < 300 else b / (2 * c ** 3)
        line = f"PG({b},{c}): mean {x.mean():.6f} (th {m:.6f}) var {x.var():.3e} (th {v:.3e})"
        if have_ref:
            from bayesbridge.random.polya_gamma import PolyaGammaDist
            from scipy.stats import ks_2samp
            r = PolyaGammaDist(3).rand_polyagamma(np.full(N, b, dtype=np.int32), np.full(N, c))
            line += f" KS p={ks_2samp(x, r).pvalue:.3f}"
        print(line)
print("== tilted stable ==")
ts = DeviceTiltedStable(ctx, 9)
for a in (1 / 32, 0.25, 0.5):
    for t in (0.01, 1.0, 10.0, 100.0, 1e4):
        x = ts.sample(a, np.full(N, t))
        line = f"TS(a={a},tilt={t}): mean {x.mean():.5e} finite {np.isfinite(x).all()}"
        if have_ref:
            from bayesbridge.random.tilted_stable import ExpTiltedStableDist
            r = ExpTiltedStableDist(4).sample(a, np.full(N // 4, t))
            line += f" ref mean {r.mean():.5e} KS p={ks_2samp(x, r).pvalue:.3f}"
        print(line)

print("== short chain (C1-like) ==")
n, p = 10000, 1000
X = rand_sparse(n, p, 0.01, True, hot=False)
beta = np.zeros(p); beta[:5] = 1.5; beta[5:10] = 1.0; beta[10:15] = 0.5
y = rng.binomial(1, 1 / (1 + np.exp(-X @ beta)))
import warnings; warnings.simplefilter('ignore')
model = bb.RegressionModel(y, X, family='logit', ctx=ctx)
bridge = bb.BayesBridge(model, bb.RegressionCoefPrior(bridge_exponent=.5))
t0 = time.time(); s, info = bridge.gibbs(n_iter=300, n_burnin=100, coef_sampler_type='cg', seed=0); t1 = time.time()
print("chain: %.1f it/s, n_cg mean %.1f, coef[:6] mean %s, tau mean %.4f" % (300 / (t1 - t0), info['_reg_coef_sampling_info']['n_cg_iter'].mean(), np.round(s['coef'][:6].mean(1), 3), s['global_scale'].mean()))
if have_ref:
    mr = ref.RegressionModel(y, X, family='logit')
    br = ref.BayesBridge(mr, ref.RegressionCoefPrior(bridge_exponent=.5))
    t0 = time.time(); sr, ir = br.gibbs(n_iter=300, n_burnin=100, coef_sampler_type='cg', seed=0); t1 = time.time()
    print("ref  : %.1f it/s, n_cg mean %.1f, coef[:6] mean %s, tau mean %.4f" % (300 / (t1 - t0), ir['_reg_coef_sampling_info']['n_cg_iter'].mean(), np.round(sr['coef'][:6].mean(1), 3), sr['global_scale'].mean()))

print("== kernel timings ==")
for (n, p, dens) in [(100000, 20000, 0.005)]:
    nnz = int(n * p * dens)
    rows = rng.integers(0, n, nnz); cols = (rng.beta(0.5, 20, nnz) * p).astype(np.int64) % p
    X = sp.csr_matrix((np.ones(nnz), (rows, cols)), shape=(n, p)); X.sum_duplicates(); X.data[:] = 1.0
    for binary in (True, False):
        for stage in (1, 0):
            ctx.set_option('spmv_stage', stage)
            D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx, pattern_only=binary)
            nz = X.nnz; bpn = 4 if binary else 12
            for what in ('dot', 'tdot', 'op'):
                ms = D.time_kernel(what, reps=20, flush_l2=True)
                ms2 = D.time_kernel(what, reps=20, flush_l2=False)
                byt = bpn * nz * (2 if what == 'op' else 1)
                print(f"n={n} p={p} nnz={nz} binary={binary} stage={stage} {what}: {ms*1e3:.1f} us cold ({byt/ms/1e6:.0f} GB/s), {ms2*1e3:.1f} us warm ({byt/ms2/1e6:.0f} GB/s)")
            del D
print("launches:", ctx.launch_count())
