"""Generates tests/golden/*.npz from the REFERENCE itself (oracle/_ref, built by oracle/build_ref.sh from
/root/reference). Run in the build container:  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_*.py) and give the GPU tests reference outputs that travel
to the GPU box, where /root/reference does not exist.
ref_saved/*.npy are the reference's own golden vectors (tests/regression_tests/saved_outputs/)."""
import os, sys, warnings
import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
warnings.simplefilter('ignore')

import bayesbridge as ref                                                     # noqa: E402
from bayesbridge.design_matrix import SparseDesignMatrix, DenseDesignMatrix  # noqa: E402
from bayesbridge.reg_coef_sampler.cg_sampler import ConjugateGradientSampler  # noqa: E402
from bayesbridge.reg_coef_sampler.reg_coef_posterior_summarizer import RegressionCoeffficientPosteriorSummarizer  # noqa: E402
from bayesbridge.random.polya_gamma import PolyaGammaDist                     # noqa: E402
from bayesbridge.random.tilted_stable import ExpTiltedStableDist              # noqa: E402
from bayesbridge.model import LinearModel, LogisticModel                      # noqa: E402


def sparse_problem(seed, n, p, density, binary):
    rs = np.random.RandomState(seed)
    X = sp.random(n, p, density=density, format='csr', random_state=rs, dtype=np.float64)
    if binary:
        X.data[:] = 1.0
    return X


def golden_random():
    rng = np.random.default_rng(1)
    shape = rng.integers(1, 5, 200).astype(np.intc)
    tilt = np.concatenate((rng.standard_normal(100) * 2, rng.standard_normal(100) * 30))
    tilt[0] = 0.
    pg = PolyaGammaDist(11).rand_polyagamma(shape, tilt)
    out = {'pg_seed': 11, 'pg_shape': shape, 'pg_tilt': tilt, 'pg_out': pg}
    for k, ce in enumerate((1 / 32, .25, .5)):
        t = np.exp(rng.standard_normal(150) * 4)
        out['ts%d_char_exp' % k] = ce
        out['ts%d_tilt' % k] = t
        out['ts%d_out' % k] = ExpTiltedStableDist(5).sample(ce, t)
    out['ts_seed'] = 5
    np.savez(os.path.join(HERE, 'random_ref.npz'), **out)


def golden_ks():
    """Reference samples for two-sample KS tests of the device samplers (superset of the grid of the
    reference's notebooks: polya_gamma/test_polyagamma.ipynb, tilted_stable/test_tilted_stable.ipynb)."""
    out = {}
    N = 4000
    grid = [(b, c) for b in (1, 2, 5) for c in (0., 0.01, 0.5, 2., 10., 50., 100.)]
    out['pg_grid'] = np.array(grid)
    pg = PolyaGammaDist(2024)
    out['pg_samples'] = np.array([pg.rand_polyagamma(np.full(N, b, dtype=np.intc), np.full(N, c)) for b, c in grid])
    tgrid = [(a, t) for a in (1 / 32, .25, .5) for t in (0.01, 1., 10., 100., 1e4)]
    out['ts_grid'] = np.array(tgrid)
    ts = ExpTiltedStableDist(2025)
    out['ts_samples'] = np.array([ts.sample(a, np.full(N, t)) for a, t in tgrid])
    np.savez_compressed(os.path.join(HERE, 'ks_ref.npz'), **out)


def golden_ks_quantiles():
    """two-sample KS at N = 1e6.  A million reference draws per grid cell cannot travel as a
    fixture (36 cells x 8 MB), so each cell's reference sample is summarised by its 4001-point quantile table
    (empirical quantiles at k/4000): the GPU test measures sup|F_device - F_reference| with F_reference interpolated
    from the table (resolution 2.5e-4, below the two-sample critical distance 3.1e-3 at n = m = 1e6)."""
    out = {}
    N, Q = 1_000_000, 4001
    probs = np.linspace(0.0, 1.0, Q)
    grid = [(b, c) for b in (1, 2, 5) for c in (0., 0.01, 0.5, 2., 10., 50., 100.)]
    out['pg_grid'] = np.array(grid)
    pg = PolyaGammaDist(4242)
    out['pg_quantiles'] = np.array([np.quantile(pg.rand_polyagamma(np.full(N, b, dtype=np.intc), np.full(N, c)), probs)
                                    for b, c in grid])
    tgrid = [(a, t) for a in (1 / 32, .25, .5) for t in (0.01, 1., 10., 100., 1e4)]
    out['ts_grid'] = np.array(tgrid)
    ts = ExpTiltedStableDist(4243)
    out['ts_quantiles'] = np.array([np.quantile(ts.sample(a, np.full(N, t)), probs) for a, t in tgrid])
    out['n_reference_draws'] = N
    np.savez_compressed(os.path.join(HERE, 'ks_quantiles_ref.npz'), **out)


def golden_design():
    out = {}
    X = sparse_problem(3, 60, 17, 0.3, False)
    out['X_dense_image'] = X.toarray()
    rng = np.random.default_rng(2)
    w, wt = rng.standard_normal(60), rng.random(60)
    out['w'], out['weight'] = w, wt
    for c in (0, 1):
        for i in (0, 1):
            D = SparseDesignMatrix(X, use_mkl=False, center_predictor=bool(c), add_intercept=bool(i))
            v = rng.standard_normal(D.shape[1])
            out['v_%d%d' % (c, i)] = v
            out['dot_%d%d' % (c, i)] = D.dot(v)
            out['tdot_%d%d' % (c, i)] = D.Tdot(w)
            out['fisher_%d%d' % (c, i)] = D.compute_fisher_info(wt, diag_only=True)
            Dd = DenseDesignMatrix(X.toarray(), center_predictor=bool(c), add_intercept=bool(i))
            out['ddot_%d%d' % (c, i)] = Dd.dot(v)
            out['dtdot_%d%d' % (c, i)] = Dd.Tdot(w)
            out['dfisher_%d%d' % (c, i)] = Dd.compute_fisher_info(wt, diag_only=True)
    np.savez(os.path.join(HERE, 'design_ref.npz'), **out)


def golden_cg():
    """Reference ConjugateGradientSampler.sample(seed=s) on seeded problems, several stopping rules."""
    out = {}
    cases = [('sparse_bin', 800, 120, 0.05), ('sparse_val', 500, 90, 0.1), ('dense', 300, 40, 1.0)]
    out['case_names'] = np.array([c[0] for c in cases])
    rules = [(1, 0.0), (5, 0.0), (20, 0.0), (500, 1e-5), (500, 1e-12)]
    out['rules'] = np.array(rules)
    for name, n, p, dens in cases:
        rng = np.random.default_rng(len(name) * 7 + 5)
        if name == 'dense':
            X = np.random.default_rng(4).standard_normal((n, p))
            D = DenseDesignMatrix(X.copy(), center_predictor=True, add_intercept=True)
            out[name + '_X'] = X
        else:
            X = sparse_problem(9, n, p, dens, name == 'sparse_bin')
            D = SparseDesignMatrix(X, use_mkl=False, center_predictor=True, add_intercept=True)
            out[name + '_indptr'], out[name + '_indices'], out[name + '_data'] = X.indptr, X.indices, X.data
            out[name + '_shape'] = np.array(X.shape)
        P = p + 1
        omega = rng.random(n) * 0.25 + 0.01
        pps = np.concatenate(([0.5], 1 / (0.1 * rng.random(p) + 1e-3)))
        z = rng.standard_normal(P)
        x0 = 0.01 * rng.standard_normal(P)
        sd = 0.5 + rng.random(P)
        out[name + '_omega'], out[name + '_pps'], out[name + '_z'] = omega, pps, z
        out[name + '_x0'], out[name + '_sd'] = x0, sd
        for k, (maxiter, atol_unit) in enumerate(rules):
            coef, info = ConjugateGradientSampler(1).sample(
                D, omega, pps, z, x0.copy(), 'prior', sd, maxiter=maxiter, atol=atol_unit * np.sqrt(P), seed=7)
            out['%s_coef_%d' % (name, k)] = coef
            out['%s_niter_%d' % (name, k)] = info['n_iter']
            out['%s_conv_%d' % (name, k)] = info['converged']
    np.savez(os.path.join(HERE, 'cg_ref.npz'), **out)


def golden_cg_c1():
    """BASELINE config 1: simulate_design(10k, 1k, binary, freq .01, seed 111), logit outcome.
    The CG inputs are the ones the reference's Gibbs sampler forms after 30 iterations of its own chain
    (reg_coef_sampler.py:60-103): omega from its PG draw, prior_prec_sqrt from its (tau, lambda), the initial guess
    and the preconditioner from its running summaries.  At this size the default rule needs 13-25 CG iterations,
    so the fixed-K cases (K = 1, 5, 10) are genuinely pre-convergence."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
    from simulate_data import simulate_design, simulate_outcome
    n, p = 10_000, 1_000
    X = simulate_design(n, p, binary_frac=1., binary_pred_freq=.01, format_='sparse', seed=111).tocsr()
    beta = np.zeros(p)
    beta[:5], beta[5:10], beta[10:15] = 1.5, 1., .5
    n_trial = np.ones(n)
    y = simulate_outcome(X, beta, intercept=0., model='logit', n_trial=n_trial, seed=1)
    if isinstance(y, tuple):
        n_success = np.asarray(y[0], dtype=float)
    else:
        n_success = np.asarray(y, dtype=float)
    model = ref.RegressionModel((n_success, n_trial), X, family='logit')
    br = ref.BayesBridge(model, ref.RegressionCoefPrior(bridge_exponent=.5))
    s, info = br.gibbs(30, 0, coef_sampler_type='cg', seed=0, params_to_save='all')
    st = info['_markov_chain_state']
    D = model.design
    Xm = D.X_main.tocsr()
    P = D.shape[1]
    omega = np.asarray(st['obs_prec'], dtype=float)
    rcs = br.reg_coef_sampler
    gscale, lscale = br.prior.adjust_scale(st['global_scale'], st['local_scale'].copy(), to='raw')
    prior_sd = np.concatenate((rcs.prior_sd_for_unshrunk, rcs.compute_prior_shrunk_scale(gscale, lscale)))
    pps = 1 / prior_sd
    z = D.Tdot(n_success - n_trial / 2)
    x0 = rcs.regcoef_summarizer.extrapolate_coef_condmean(gscale, lscale)
    sd = rcs.regcoef_summarizer.estimate_coef_precond_scale_sd().copy()
    rules = [(1, 0.0), (5, 0.0), (10, 0.0), (500, 1e-5), (500, 1e-12)]
    out = {'indptr': Xm.indptr, 'indices': Xm.indices, 'shape': np.array(Xm.shape), 'n_success': n_success,
           'omega': omega, 'pps': pps, 'z': z, 'x0': x0, 'sd': sd, 'rules': np.array(rules),
           'chain_n_cg': info['_reg_coef_sampling_info']['n_cg_iter']}
    assert np.all(Xm.data == 1.0)
    for k, (maxiter, atol_unit) in enumerate(rules):
        coef, cinfo = ConjugateGradientSampler(1).sample(
            D, omega, pps, z, x0.copy(), 'prior', sd, maxiter=maxiter, atol=atol_unit * np.sqrt(P), seed=7)
        out['coef_%d' % k], out['niter_%d' % k], out['conv_%d' % k] = coef, cinfo['n_iter'], cinfo['converged']
        print('cg_c1 rule', (maxiter, atol_unit), 'n_iter', cinfo['n_iter'], 'converged', cinfo['converged'])
    np.savez_compressed(os.path.join(HERE, 'cg_c1_ref.npz'), **out)


def golden_summarizer():
    rng = np.random.default_rng(8)
    S = RegressionCoeffficientPosteriorSummarizer(12, 2, 1.5)
    x0s, sds, coefs, gs, ls = [], [], [], [], []
    for it in range(6):
        g, l = float(rng.random() + 0.1), rng.random(10) + 0.2
        x0s.append(S.extrapolate_coef_condmean(g, l))
        sds.append(S.estimate_coef_precond_scale_sd().copy())
        c = rng.standard_normal(12)
        S.update(c, g, l)
        coefs.append(c); gs.append(g); ls.append(l)
    np.savez(os.path.join(HERE, 'summarizer_ref.npz'), x0=np.array(x0s), sd=np.array(sds), coef=np.array(coefs),
             gscale=np.array(gs), lscale=np.array(ls))


def test_gibb_data(model, fmt):
    """Data of the reference's tests/regression_tests/test_gibb.py:62-90."""
    np.random.seed(1)
    n, p = 100, 50
    beta = np.zeros(p)
    beta[:4] = 1
    beta[4:15] = 2 ** - np.linspace(0.0, 5, 11)
    X = np.random.randn(n, p)
    if model == 'linear':
        outcome = LinearModel.simulate_outcome(X, beta, 2)
    else:
        n_trial = np.ones(n, dtype=np.int32)
        outcome = (LogisticModel.simulate_outcome(n_trial, X, beta), n_trial)
    return outcome, (sp.csr_matrix(X) if fmt == 'sparse' else X)


def golden_chain():
    out = {}
    for fam, fmt in (('linear', 'dense'), ('logit', 'sparse')):
        outcome, X = test_gibb_data(fam, fmt)
        prior = ref.RegressionCoefPrior(sd_for_intercept=2., regularizing_slab_size=1., bridge_exponent=0.25)
        br = ref.BayesBridge(ref.RegressionModel(outcome, X, fam), prior)
        s, info = br.gibbs(10, 0, init={'global_scale': 0.1, 'local_scale': np.ones(50)},
                           coef_sampler_type='cg', seed=0, params_to_save='all')
        out[fam + '_coef'] = s['coef']
        out[fam + '_gscale'] = s['global_scale']
        out[fam + '_n_cg'] = info['_reg_coef_sampling_info']['n_cg_iter']
        out[fam + '_X'] = X.toarray() if fmt == 'sparse' else X
        if fam == 'logit':
            out['logit_n_success'], out['logit_n_trial'] = outcome
        else:
            out['linear_y'] = outcome
    # the 'cholesky' combo of the reference's regression test (test_gibb.py:12-13: logit, dense X)
    outcome, X = test_gibb_data('logit', 'dense')
    prior = ref.RegressionCoefPrior(sd_for_intercept=2., regularizing_slab_size=1., bridge_exponent=0.25)
    br = ref.BayesBridge(ref.RegressionModel(outcome, X, 'logit'), prior)
    s, info = br.gibbs(10, 0, init={'global_scale': 0.1, 'local_scale': np.ones(50)},
                       coef_sampler_type='cholesky', seed=0, params_to_save='all')
    out['logitchol_coef'], out['logitchol_gscale'] = s['coef'], s['global_scale']
    np.savez(os.path.join(HERE, 'chain_ref.npz'), **out)


def golden_cholesky():
    """Reference outputs of the direct sampler's pieces: compute_fisher_info(weight) (full matrix) of the dense and the
    sparse class, and generate_gaussian_with_weight with numpy's global stream seeded (direct_gaussian_sampler.py)."""
    from bayesbridge.reg_coef_sampler.direct_gaussian_sampler import generate_gaussian_with_weight
    out = {}
    rng = np.random.default_rng(21)
    n, p = 400, 37
    Xd = rng.standard_normal((n, p))
    Xs = sparse_problem(5, n, p, 0.15, False)
    out['Xd'], out['Xs_dense_image'] = Xd, Xs.toarray()
    w = rng.random(n) * 0.25 + 0.01
    out['weight'] = w
    for c in (0, 1):
        for i in (0, 1):
            Dd = DenseDesignMatrix(Xd.copy(), center_predictor=bool(c), add_intercept=bool(i))
            Ds = SparseDesignMatrix(Xs, use_mkl=False, center_predictor=bool(c), add_intercept=bool(i))
            out['fisher_dense_%d%d' % (c, i)] = Dd.compute_fisher_info(w)
            out['fisher_sparse_%d%d' % (c, i)] = Ds.compute_fisher_info(w)
    P = p + 1
    pps = np.concatenate(([0.5], 1 / (0.1 * rng.random(p) + 1e-2)))
    z = rng.standard_normal(P)
    out['pps'], out['z'] = pps, z
    for name, D in (('dense', DenseDesignMatrix(Xd.copy(), center_predictor=True, add_intercept=True)),
                    ('sparse', SparseDesignMatrix(Xs, use_mkl=False, center_predictor=True, add_intercept=True))):
        np.random.seed(13)
        out['draw_' + name] = generate_gaussian_with_weight(D, w, pps, z)
    np.savez(os.path.join(HERE, 'cholesky_ref.npz'), **out)


def golden_posterior():
    """A long reference chain on the C1-like problem of tests/test_gpu_gibbs.py::_c1_like."""
    rs = np.random.RandomState(0)
    n, p = 4000, 300
    X = sp.random(n, p, density=0.03, format='csr', random_state=rs, dtype=np.float64)
    X.data[:] = 1.0
    beta = np.zeros(p)
    beta[:5], beta[5:10] = 1.5, -1.0
    y = rs.binomial(1, 1 / (1 + np.exp(-(X @ beta - 0.5))))
    br = ref.BayesBridge(ref.RegressionModel(y, X, family='logit'), ref.RegressionCoefPrior(bridge_exponent=.5))
    s, info = br.gibbs(n_iter=2500, n_burnin=500, coef_sampler_type='cg', seed=0)
    # thinned marginals (every 20th of the 2000 kept draws) of the intercept, the ten signal coefficients, five null ones,
    # log tau and the log-posterior: the two-sample KS tests of test_device_rng_chain_matches_reference_posterior
    thin_idx = np.array(list(range(0, 11)) + [20, 50, 100, 200, 300])
    np.savez(os.path.join(HERE, 'posterior_ref.npz'), y=y, coef_mean=s['coef'].mean(1), coef_sd=s['coef'].std(1),
             log_gscale_mean=np.log(s['global_scale']).mean(), logp_mean=s['logp'].mean(),
             n_cg_mean=info['_reg_coef_sampling_info']['n_cg_iter'].mean(),
             thin_idx=thin_idx, thin_coef=s['coef'][thin_idx][:, ::20], thin_log_gscale=np.log(s['global_scale'][::20]),
             thin_logp=s['logp'][::20])


if __name__ == '__main__':
    only = sys.argv[1:]
    todo = [golden_random, golden_ks, golden_ks_quantiles, golden_design, golden_cg, golden_cg_c1, golden_summarizer, golden_chain,
            golden_cholesky, golden_posterior]
    for fn in todo:
        if not only or fn.__name__ in only:
            fn()
    print('golden fixtures written to', HERE)
