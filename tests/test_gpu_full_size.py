"""GPU: the BENCHMARKED configuration (BASELINE config 4: binary sparse 1M x 100k, 0.1 %, bench.py's generator, nnz ~ 1e8)
checked entry by entry against the CPU checker:

* dot / Tdot / fisher-diag vs scipy on the full matrix (scipy does a 1e8-nnz product in ~0.2 s);
* one coefficient draw `bb_cg_sample` with injected omega and noise vs `oracle.cg_oracle.cg_sample` on the same inputs,
  taken from the state the chain has reached after a few Gibbs iterations (so the system is as hard as the ones the
  bench solves): the default stopping rule (1e-5 sqrt(P), reg_coef_sampler.py:95) must stop after the same number of
  iterations (a few more or fewer where rounding moves the threshold crossing: see the comment in the test) and agree
  to 1e-7 (1e-5 when the iteration counts differ); both sides converged to 1e-12 sqrt(P) must agree to the north-star's 1e-8.
"""
import numpy as np
import pytest

from conftest import record_achieved
from oracle import cg_oracle as co

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope='module')
def c4(ctx):
    import bench
    import bayesbridge_b200 as bb
    from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix
    n, p, dens = bench.WORKLOADS['C4']
    X, y = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
    D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx)
    Xm = D.X_main                       # the host image after remove_intercept_indicator (abstract_matrix.py:93-107)
    O = co.DesignOracle(Xm, True, True)
    return dict(X=Xm, y=y, D=D, O=O, bb=bb)


def test_c4_products_vs_scipy(ctx, c4):
    D, O = c4['D'], c4['O']
    n, P = D.shape
    assert D.is_binary and n == 1_000_000 and D.nnz > 9.9e7
    rng = np.random.default_rng(0)
    v, w, wt = rng.standard_normal(P), rng.standard_normal(n), rng.random(n)
    for name, got, ref, bound in (('dot', D.dot(v), O.dot(v), 1e-13), ('Tdot', D.Tdot(w), O.Tdot(w), 1e-13),
                                  ('fisher_diag', D.compute_fisher_info(wt, diag_only=True), O.fisher_diag(wt), 1e-12)):
        err = relerr(got, ref)
        worst = float(np.abs(got - ref).max() / np.abs(ref).max())
        record_achieved('c4_products_vs_scipy', name, err, bound, max_abs_over_max=worst)
        assert err <= bound, (name, err)
        assert worst <= 10 * bound, (name, worst)


def test_c4_cg_sample_vs_oracle(ctx, c4):
    from bayesbridge_b200.reg_coef_sampler import ConjugateGradientSampler
    bb, D, O, y = c4['bb'], c4['D'], c4['O'], c4['y']
    n, P = D.shape
    # a few Gibbs iterations of the real sampler give omega, tau, lambda and the running summaries of a live chain
    model = bb.RegressionModel(y, D, family='logit')
    bridge = bb.BayesBridge(model, bb.RegressionCoefPrior(bridge_exponent=.5))
    _, info = bridge.gibbs(n_iter=3, n_burnin=0, coef_sampler_type='cg', seed=0, params_to_save=('global_scale',))
    st = info['_markov_chain_state']
    omega = np.ascontiguousarray(st['obs_prec'], dtype=float)
    gscale, lscale = bridge.prior.adjust_scale(st['global_scale'], np.array(st['local_scale'], copy=True), to='raw')
    rcs = bridge.reg_coef_sampler
    prior_sd = np.concatenate((rcs.prior_sd_for_unshrunk, rcs.compute_prior_shrunk_scale(gscale, lscale)))
    pps = 1 / prior_sd
    z = O.Tdot(y - 0.5)                                   # kappa = n_success - n_trial / 2
    x0 = np.asarray(st['coef'], dtype=float).copy()       # a warm start of the size the summarizer gives
    sd = np.ones(P)
    s = co.precond_scale_prior(pps, 1, sd)
    np.random.seed(11)
    e1, e2 = np.random.randn(n), np.random.randn(P)
    for atol_unit, bound in ((1e-5, 1e-7), (1e-12, 1e-8)):
        atol = atol_unit * np.sqrt(P)
        ref, rinfo = co.cg_sample(O, omega, pps, z, x0, s, 500, atol, e1, e2)
        coef, cinfo = ConjugateGradientSampler(1).sample(D, omega, pps, z, x0.copy(), 'prior', sd, maxiter=500, atol=atol, seed=11)
        err = relerr(coef, ref)
        if cinfo['n_iter'] != rinfo['n_iter']:
            # ~100 iterations into an ill-conditioned solve, finite-precision CG has lost orthogonality and the residual
            # norms of two correct implementations differ by tens of per cent (residual spikes where p.q nearly
            # vanishes: scripts/diag_niter.py measures a factor 17 between this library's own kernel variants on the
            # config-1 fixture), so the iteration at which ||r|| crosses the threshold moves by a few iterations.  What
            # the stopping rule promises is a solution within the tolerance: the two stopped solutions may differ by a
            # few late CG steps at the default rule, and must still agree to 1e-8 once both are converged.
            assert abs(cinfo['n_iter'] - rinfo['n_iter']) <= max(2, rinfo['n_iter'] // 20), (cinfo['n_iter'], rinfo['n_iter'])
            if atol_unit == 1e-5:
                bound = 1e-5
        record_achieved('c4_cg_sample_vs_oracle', atol_unit, err, bound, n_iter=cinfo['n_iter'], n_iter_oracle=rinfo['n_iter'])
        assert cinfo['converged'] and rinfo['converged']
        assert err <= bound, (atol_unit, err)
