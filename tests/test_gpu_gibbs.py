"""GPU: the drop-in Gibbs sampler.
(i)  compatibility mode -- host RNG for the CG noise (options={'noise': 'host'}) and the oracle's
     bit-exact ports of the reference's PG / tilted-stable streams plugged into BasicRandom -- replays the
     reference's regression test (tests/regression_tests/test_gibb.py, 'cg' combos) and the cupy parity
     test template (tests/gpu_tests/test_gibbs.py): chain vs the reference chain and its saved vectors;
(ii) device-RNG mode -- posterior summaries agree with a reference chain within Monte-Carlo error."""
import os
import numpy as np
import scipy.sparse as sp
import pytest

from conftest import golden, GOLDEN, import_reference
from oracle.rand_port import PolyaGammaPort, TiltedStablePort

pytestmark = pytest.mark.gpu


def _bb():
    import bayesbridge_b200 as bb
    return bb


def _test_gibb_problem(family):
    g = golden('chain_ref.npz')
    X = g[family + '_X']
    if family == 'logit':
        return (g['logit_n_success'], g['logit_n_trial']), sp.csr_matrix(X), g
    return g['linear_y'], X, g


def _compat_bridge(family, ctx):
    bb = _bb()
    outcome, X, g = _test_gibb_problem(family)
    prior = bb.RegressionCoefPrior(sd_for_intercept=2., regularizing_slab_size=1., bridge_exponent=0.25)
    bridge = bb.BayesBridge(bb.RegressionModel(outcome, X, family, ctx=ctx), prior)
    bridge.rg.pg, bridge.rg.ts = PolyaGammaPort(), TiltedStablePort()     # the reference's streams
    return bridge, g


@pytest.mark.parametrize('family', ['linear', 'logit'])
def test_compat_chain_reproduces_reference(ctx, family):
    bridge, g = _compat_bridge(family, ctx)
    samples, info = bridge.gibbs(10, 0, init={'global_scale': 0.1, 'local_scale': np.ones(50)},
                                 coef_sampler_type='cg', seed=0, params_to_save='all',
                                 options={'noise': 'host', 'init_optimizer': 'scipy'})
    saved = np.load(os.path.join(GOLDEN, 'ref_saved', family + '_cg_samples.npy'))
    assert np.allclose(samples['coef'][:, -1], saved, rtol=.001, atol=10e-6)       # reference's own criterion
    # every sample vs the reference chain; the cupy-parity test of the reference uses atol 1e-5
    # (gpu_tests/test_gibbs.py:44); each CG solve is only converged to ~1e-5, hence the factor 5
    assert np.allclose(samples['coef'], g[family + '_coef'], rtol=0, atol=5e-5)
    assert np.allclose(samples['global_scale'], g[family + '_gscale'], rtol=1e-4)
    n_cg = info['_reg_coef_sampling_info']['n_cg_iter']
    assert np.max(np.abs(n_cg - g[family + '_n_cg'])) <= 2     # the stopping test sits at the tolerance boundary


def test_compat_cholesky_chain_reproduces_reference(ctx):
    """The 'cholesky' combo of the reference's regression test (test_gibb.py:12-13: logit, dense X, 10 iterations, with
    and without a restart in the middle) through the device's direct sampler."""
    bb = _bb()
    outcome, X, g = _test_gibb_problem('logit')
    X = X.toarray()
    prior = bb.RegressionCoefPrior(sd_for_intercept=2., regularizing_slab_size=1., bridge_exponent=0.25)
    init = {'global_scale': 0.1, 'local_scale': np.ones(50)}

    def bridge():
        b = bb.BayesBridge(bb.RegressionModel(outcome, X, 'logit', ctx=ctx), prior)
        b.rg.pg, b.rg.ts = PolyaGammaPort(), TiltedStablePort()
        return b
    samples, info = bridge().gibbs(10, 0, init=init, coef_sampler_type='cholesky', seed=0, params_to_save='all',
                                   options={'init_optimizer': 'scipy'})
    assert info['coef_sampler_type'] == 'cholesky'
    saved = np.load(os.path.join(GOLDEN, 'ref_saved', 'logit_cholesky_samples.npy'))
    assert np.allclose(samples['coef'][:, -1], saved, rtol=.001, atol=10e-6)       # the reference's own criterion
    assert np.allclose(samples['coef'], g['logitchol_coef'], rtol=0, atol=1e-8)    # the reference chain run here
    assert np.allclose(samples['global_scale'], g['logitchol_gscale'], rtol=1e-9)
    s1, i1 = bridge().gibbs(5, 0, init=init, coef_sampler_type='cholesky', seed=0, params_to_save='all',
                            options={'init_optimizer': 'scipy'})
    s2, _ = bridge().gibbs_resume(i1, 5, merge=True, prev_samples=s1)
    assert np.allclose(s2['coef'], g['logitchol_coef'], rtol=0, atol=1e-8)


def test_compat_chain_resume_equals_uninterrupted(ctx):
    bridge, g = _compat_bridge('logit', ctx)
    init = {'global_scale': 0.1, 'local_scale': np.ones(50)}
    s1, i1 = bridge.gibbs(5, 0, init=init, coef_sampler_type='cg', seed=0, params_to_save='all',
                          options={'noise': 'host', 'init_optimizer': 'scipy'})
    bridge2, _ = _compat_bridge('logit', ctx)
    s2, i2 = bridge2.gibbs_resume(i1, 5, merge=True, prev_samples=s1)
    assert s2['coef'].shape == (51, 10) and i2['n_iter'] == 10
    assert np.allclose(s2['coef'], g['logit_coef'], rtol=0, atol=5e-5)


def test_device_chain_resume_is_exact(ctx):
    """Device RNG: the chain state (coef, scales, omega) + Philox (seed, offset) + numpy state resumes bit-exactly."""
    bb = _bb()
    outcome, X, _ = _test_gibb_problem('logit')
    prior = bb.RegressionCoefPrior(sd_for_intercept=2., regularizing_slab_size=1., bridge_exponent=0.5)
    init = {'global_scale': 0.1, 'local_scale': np.ones(50)}
    full, _ = bb.BayesBridge(bb.RegressionModel(outcome, X, 'logit', ctx=ctx), prior).gibbs(
        12, 0, init=init, coef_sampler_type='cg', seed=3, params_to_save='all')
    b1 = bb.BayesBridge(bb.RegressionModel(outcome, X, 'logit', ctx=ctx), prior)
    s1, i1 = b1.gibbs(6, 0, init=init, coef_sampler_type='cg', seed=3, params_to_save='all')
    b2 = bb.BayesBridge(bb.RegressionModel(outcome, X, 'logit', ctx=ctx), prior)
    s2, _ = b2.gibbs_resume(i1, 6, merge=True, prev_samples=s1)
    assert np.array_equal(s2['coef'], full['coef'])
    assert np.array_equal(s2['obs_prec'], full['obs_prec'])


def _c1_like(n=4000, p=300, seed=0):
    rs = np.random.RandomState(seed)
    X = sp.random(n, p, density=0.03, format='csr', random_state=rs, dtype=np.float64)
    X.data[:] = 1.0
    beta = np.zeros(p)
    beta[:5], beta[5:10] = 1.5, -1.0
    y = rs.binomial(1, 1 / (1 + np.exp(-(X @ beta - 0.5))))
    return y, X, beta


def test_device_rng_chain_matches_reference_posterior(ctx):
    """Posterior means of a device-RNG chain vs a reference chain (different random streams): agreement within
    Monte-Carlo error. Uses the reference's saved chain summary in golden/posterior_ref.npz."""
    bb = _bb()
    g = golden('posterior_ref.npz')
    y, X, beta = _c1_like()
    assert np.array_equal(y, g['y'])
    model = bb.RegressionModel(y, X, family='logit', ctx=ctx)
    bridge = bb.BayesBridge(model, bb.RegressionCoefPrior(bridge_exponent=.5))
    samples, info = bridge.gibbs(n_iter=1500, n_burnin=500, coef_sampler_type='cg', seed=1)
    assert info['options']['coef_sampler_type'] == 'cg'
    mean, sd = samples['coef'].mean(1), samples['coef'].std(1)
    ref_mean, ref_sd = g['coef_mean'], g['coef_sd']
    # MCMC standard errors: allow 6 se with an effective sample size of ~ N/10 on both sides, plus CG tolerance
    se = np.sqrt(sd ** 2 + ref_sd ** 2) / np.sqrt(100)
    assert np.all(np.abs(mean - ref_mean) < 6 * se + 1e-3)
    big = np.abs(ref_mean) > 0.5
    assert big.sum() >= 8 and np.allclose(mean[big], ref_mean[big], rtol=0.08)        # achieved on B200: 0.027
    assert np.log(samples['global_scale']).mean() == pytest.approx(float(g['log_gscale_mean']), abs=0.15)   # achieved: 0.025
    assert samples['logp'].mean() == pytest.approx(float(g['logp_mean']), rel=0.01)                        # achieved: 0.0036
    n_cg = info['_reg_coef_sampling_info']['n_cg_iter']
    assert n_cg.mean() == pytest.approx(float(g['n_cg_mean']), rel=0.06)                                   # achieved: 0.015
    from conftest import record_achieved as _rec
    _rec('device_rng_chain', 'max |mean - ref| / se over 301 coefficients', float(np.max(np.abs(mean - ref_mean) / (se + 1e-3 / 6))), 6.0)
    _rec('device_rng_chain', 'max rel diff of the big coefficients', float(np.max(np.abs(mean[big] / ref_mean[big] - 1))), 0.08)
    _rec('device_rng_chain', '|mean log tau - ref|', abs(float(np.log(samples['global_scale']).mean()) - float(g['log_gscale_mean'])), 0.15)
    _rec('device_rng_chain', 'rel diff of mean logp', abs(float(samples['logp'].mean()) / float(g['logp_mean']) - 1), 0.01)
    _rec('device_rng_chain', 'rel diff of mean n_cg_iter', abs(float(n_cg.mean()) / float(g['n_cg_mean']) - 1), 0.06)
    # Two-sample Kolmogorov-Smirnov on thinned marginals: 100 thinned draws of the reference chain
    # (every 20th of 2000) against 100 of this chain (every 10th of 1000) for the intercept, the ten signal coefficients,
    # five null coefficients, log tau and the log-posterior.  For n = m = 100 the critical distance is 0.276 at alpha = 1e-3
    # and 0.349 at alpha = 1e-5; two REFERENCE chains with different seeds differ by 0.06 ... 0.17 on these marginals, median
    # 0.12 (measured when the fixture was made).  Every marginal must stay below the 1e-5 distance and the median over the
    # 18 marginals below 0.16.  (The chain is seeded and the kernels are deterministic, so the outcome does not flicker;
    # the bounds leave room for a change of summation order in a later build.)
    from scipy.stats import ks_2samp
    from conftest import record_achieved
    dist = {}
    for row, j in enumerate(g['thin_idx']):
        dist['coef[%d]' % int(j)] = ks_2samp(samples['coef'][int(j), ::10], g['thin_coef'][row]).statistic
    dist['log_tau'] = ks_2samp(np.log(samples['global_scale'][::10]), g['thin_log_gscale']).statistic
    dist['logp'] = ks_2samp(samples['logp'][::10], g['thin_logp']).statistic
    for case, d in dist.items():
        record_achieved('device_rng_chain_ks', case, d, 0.349)
    record_achieved('device_rng_chain_ks', 'median over 18 marginals', float(np.median(list(dist.values()))), 0.16)
    assert max(dist.values()) < 0.349, dist
    assert np.median(list(dist.values())) < 0.16, dist


def test_linear_dense_device_chain_recovers_signal(ctx):
    bb = _bb()
    rng = np.random.default_rng(0)
    n, p = 3000, 120
    X = rng.standard_normal((n, p))
    beta = np.zeros(p); beta[:6] = (2., -2., 1., -1., .5, -.5)
    y = 1.0 + X @ beta + rng.standard_normal(n)
    bridge = bb.BayesBridge(bb.RegressionModel(y, X, 'linear', ctx=ctx), bb.RegressionCoefPrior(bridge_exponent=.5))
    s, info = bridge.gibbs(400, 100, coef_sampler_type='cg', seed=0, params_to_save='all')
    m = s['coef'].mean(1)
    assert m[0] == pytest.approx(1.0, abs=0.1)
    assert np.allclose(m[1:7], beta[:6], atol=0.1)
    assert np.abs(m[7:]).max() < 0.1
    assert s['obs_prec'].mean() == pytest.approx(1.0, rel=0.15)


def test_api_contract(ctx):
    bb = _bb()
    y, X, _ = _c1_like(600, 40)
    model = bb.RegressionModel(y, X, family='logit', ctx=ctx)
    assert model.design.use_gpu and model.design.is_sparse and not model.design.use_cupy
    bridge = bb.BayesBridge(model, bb.RegressionCoefPrior())
    with pytest.raises(ValueError):                      # as the reference does for cupy matrices
        bridge.gibbs(n_iter=1, coef_sampler_type='hmc')
    s, info = bridge.gibbs(n_iter=2, coef_sampler_type='cholesky', seed=1)     # the direct sampler, also on a sparse design
    assert info['coef_sampler_type'] == 'cholesky' and np.all(np.isfinite(s['coef']))
    s, info = bridge.gibbs(n_iter=3, init={'coef': np.ones(model.n_pred)}, seed=1)   # gpu_tests/test_gibbs.py:32-44
    assert info['options']['coef_sampler_type'] == 'cg'
    assert s['coef'].shape == (41, 3) and s['logp'].shape == (3,)
    for key in ('_markov_chain_state', '_random_gen_state', '_reg_coef_sampler_state', 'runtime'):
        assert key in info
    assert info['_markov_chain_state']['obs_prec'].shape == (600,)
    with pytest.raises(NotImplementedError):
        bb.RegressionModel((y, y), X, family='cox', ctx=ctx)
    with pytest.raises(ValueError):
        bb.RegressionModel(y[:-1], X, family='logit', ctx=ctx)


def test_resident_state_path_equals_host_path(ctx, monkeypatch):
    """The P-side Gibbs state on the device (bb_state_* / bb_cg_sample_resident / bb_local_scale_resident) mirrors the
    host arithmetic: with the same seeds the first coefficient draw is bit-identical to the host path and the first
    few iterations agree to CG-amplified round-off (the only difference is the summation order of sum|beta|^alpha,
    which enters through the Gamma draw of tau); the saved local scales and the summaries are equal too."""
    bb = _bb()
    y, X, _ = _c1_like(3000, 200, seed=4)
    prior = bb.RegressionCoefPrior(bridge_exponent=.5, sd_for_intercept=2., regularizing_slab_size=3.)
    runs = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('BB_RESIDENT_STATE', flag)
        bridge = bb.BayesBridge(bb.RegressionModel(y, X, 'logit', ctx=ctx), prior)
        s, info = bridge.gibbs(6, 0, coef_sampler_type='cg', seed=11, params_to_save='all')
        runs[flag] = (s, info)
    host, dev = runs['0'][0], runs['1'][0]
    assert np.array_equal(dev['coef'][:, 0], host['coef'][:, 0])
    assert np.array_equal(dev['obs_prec'][:, 0], host['obs_prec'][:, 0])
    assert dev['global_scale'][0] == pytest.approx(host['global_scale'][0], rel=1e-12)
    assert np.allclose(dev['local_scale'][:, 0], host['local_scale'][:, 0], rtol=1e-10)
    assert dev['logp'][0] == pytest.approx(host['logp'][0], rel=1e-10)
    assert np.allclose(dev['coef'][:, 1], host['coef'][:, 1], rtol=0, atol=1e-6)
    sh = runs['0'][1]['_reg_coef_sampler_state']['regcoef_summarizer'].coef_scaled_summarizer
    sd = runs['1'][1]['_reg_coef_sampler_state']['regcoef_summarizer'].coef_scaled_summarizer
    assert sd.n_averaged == sh.n_averaged == 6
    assert np.allclose(sd.stats['mean'], sh.stats['mean'], rtol=0, atol=1e-4)
    n_cg_h = runs['0'][1]['_reg_coef_sampling_info']['n_cg_iter']
    n_cg_d = runs['1'][1]['_reg_coef_sampling_info']['n_cg_iter']
    assert n_cg_h[0] == n_cg_d[0] and np.max(np.abs(n_cg_h - n_cg_d)) <= 3


def test_resident_state_linear_model(ctx, monkeypatch):
    bb = _bb()
    rng = np.random.default_rng(2)
    n, p = 1500, 60
    X = rng.standard_normal((n, p))
    y = 0.5 + X[:, :3] @ np.array([1.5, -1., .5]) + rng.standard_normal(n)
    out = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('BB_RESIDENT_STATE', flag)
        bridge = bb.BayesBridge(bb.RegressionModel(y, X, 'linear', ctx=ctx), bb.RegressionCoefPrior(bridge_exponent=.5))
        out[flag], _ = bridge.gibbs(4, 0, coef_sampler_type='cg', seed=5, params_to_save='all')
    assert np.array_equal(out['1']['coef'][:, 0], out['0']['coef'][:, 0])
    assert np.allclose(out['1']['coef'][:, 1], out['0']['coef'][:, 1], atol=1e-6)
    assert out['1']['obs_prec'][0] == pytest.approx(out['0']['obs_prec'][0], rel=1e-12)


@pytest.mark.parametrize('family', ['logit', 'linear'])
def test_device_mode_search_finds_the_optimum_scipy_finds(ctx, family):
    """Chain initialisation (reg_coef_sampler.py:281-358): the L-BFGS search inside libbbgpu (bb_mode_search) and scipy's
    L-BFGS-B driven through the same device likelihood reach the same conditional posterior mode."""
    bb = _bb()
    from bayesbridge_b200.reg_coef_sampler import SparseRegressionCoefficientSampler
    rs = np.random.RandomState(4)
    n, p = 20000, 600
    X = sp.random(n, p, density=0.02, format='csr', random_state=rs, dtype=np.float64)
    X.data[:] = 1.0
    beta = np.zeros(p); beta[:8] = 1.2
    eta = X @ beta - 0.7
    if family == 'logit':
        outcome, obs_prec = rs.binomial(1, 1 / (1 + np.exp(-eta))), None
    else:
        outcome, obs_prec = eta + rs.randn(n), 0.8
    model = bb.RegressionModel(outcome, X, family=family, ctx=ctx)
    P = model.n_pred
    lscale, gscale = np.exp(rs.randn(P - 1)), 0.05
    found = {}
    for opt in ('scipy', 'device'):
        S = SparseRegressionCoefficientSampler(P, np.array([float('inf')]), 'cg')
        S.init_optimizer = opt
        coef0 = np.zeros(P); coef0[0] = model.calc_intercept_mle()
        coef, info = S.search_mode(coef0, lscale, gscale, obs_prec, model)
        assert info['is_success'], (opt, info)
        scale, prior_prec = S.compute_preconditioning_scale(gscale, lscale, np.ones(P), S.prior_sd_for_unshrunk)
        args = (obs_prec,) if family == 'linear' else ()
        ll, grad = model.compute_loglik_and_gradient(coef, *args)
        theta = coef / scale
        F = -ll + 0.5 * np.sum(prior_prec * theta ** 2)
        gmax = np.abs(-scale * grad + prior_prec * theta).max()
        found[opt] = (coef, F, gmax, info)
    (c_s, F_s, g_s, i_s), (c_d, F_d, g_d, i_d) = found['scipy'], found['device']
    gtol = 1e-6 / np.sqrt(P)
    assert g_d <= 20 * gtol or i_d['n_iter'] > 0 and abs(F_d - F_s) <= 1e-9 * abs(F_s), (g_d, gtol)
    assert F_d <= F_s + 1e-9 * abs(F_s), (F_d, F_s)
    assert np.linalg.norm(c_d - c_s) <= 1e-4 * np.linalg.norm(c_s), np.linalg.norm(c_d - c_s) / np.linalg.norm(c_s)
    print('mode search: scipy', i_s['n_iter'], 'iterations, device', i_d['n_iter'], 'iterations; |grad|_inf', g_s, g_d)
