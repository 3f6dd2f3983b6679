"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's coefficient update.

Follows, function by function:
  design algebra      design_matrix/sparse_matrix.py:68-129,164-177 ; dense_matrix.py:20-58
  CG sampler          reg_coef_sampler/cg_sampler.py:20-151
  CG loop             scipy.sparse.linalg.cg (third party, unpinned by the reference; restated from
                      scipy 1.18.1 _isolve/iterative.py:cg -- test ||r||<atol first, rho=r.r, p=r+beta p,
                      q=Ap, alpha=rho/(p.q), x+=alpha p, r-=alpha q, callback; maxiter exhaustion -> info=maxiter)
  summarizer          reg_coef_sampler/reg_coef_posterior_summarizer.py:12-41,88-124
  coefficient update  reg_coef_sampler/reg_coef_sampler.py:60-103,194-201
  Gibbs loop          bayesbridge.py:109-277,355-511 (only what the 'cg' path touches)
Pinned in tests/test_oracle_cg.py against fixtures produced by the reference (tests/golden/)."""
import math
import numpy as np
import scipy.sparse as sparse


# ---- design matrix ---------------------------------------------------------------------------
class DesignOracle:
    """X_main (scipy CSR or dense ndarray) + implicit intercept + implicit centring."""

    def __init__(self, X, center_predictor=False, add_intercept=True):
        self.sparse = sparse.issparse(X)
        if self.sparse:
            self.X = X.tocsr()
            mean = np.squeeze(np.array(X.mean(axis=0))).reshape(-1)
        else:
            self.X = np.asarray(X, dtype=float)
            mean = self.X.mean(axis=0)
        self.c = mean if center_predictor else np.zeros(self.X.shape[1])
        self.centered = center_predictor
        self.icpt = int(add_intercept)
        self.shape = (self.X.shape[0], self.X.shape[1] + self.icpt)

    def _materialised(self):
        # dense_matrix.py:20-25: the reference's dense class stores [1, X - mean] explicitly
        if not hasattr(self, '_A'):
            A = self.X - self.c[None, :] if self.centered else self.X.copy()
            if self.icpt:
                A = np.hstack((np.ones((A.shape[0], 1)), A))
            self._A = A
        return self._A

    def dot(self, v):
        if not self.sparse:
            return self._materialised().dot(v)
        v0 = v[0] if self.icpt else 0.
        v1 = v[self.icpt:]
        out = v0 + self.X.dot(v1)
        out -= np.inner(self.c, v1)
        return out

    def Tdot(self, w):
        if not self.sparse:
            return self._materialised().T.dot(w)
        t = self.X.T.dot(w)
        t = t - np.sum(w) * self.c
        if self.icpt:
            t = np.concatenate(([np.sum(w)], t))
        return t

    def fisher_diag(self, weight):
        if not self.sparse:
            return np.sum(weight[:, np.newaxis] * self._materialised() ** 2, 0)
        if self.sparse:
            X2 = self.X.multiply(self.X)
            d = np.asarray(X2.T.dot(weight)).ravel()
            wx = np.asarray(self.X.T.dot(weight)).ravel()
        else:
            d = (self.X ** 2).T.dot(weight)
            wx = self.X.T.dot(weight)
        if self.centered:
            d = d - 2 * self.c * wx + np.sum(weight) * self.c ** 2
        if self.icpt:
            d = np.concatenate(([np.sum(weight)], d))
        return d

    def toarray(self):
        A = (self.X.toarray() if self.sparse else self.X) - self.c[None, :]
        if self.icpt:
            A = np.hstack((np.ones((A.shape[0], 1)), A))
        return A


# ---- CG ----------------------------------------------------------------------------------------
def cg_loop(matvec, b, x0, maxiter, atol):
    """scipy.sparse.linalg.cg with M = identity, rtol folded into atol. Returns x, info, n_iter."""
    x = x0.copy()
    if np.linalg.norm(b) == 0:
        return b.copy(), 0, 0
    r = b - matvec(x) if x.any() else b.copy()
    rho_prev, p, n_iter = None, None, 0
    for iteration in range(maxiter):
        if np.linalg.norm(r) < atol:
            return x, 0, n_iter
        rho = np.dot(r, r)
        if iteration > 0:
            p *= rho / rho_prev
            p += r
        else:
            p = r.copy()
        q = matvec(p)
        alpha = rho / np.dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
        n_iter += 1
    return x, maxiter, n_iter


def precond_scale_prior(prior_prec_sqrt, n_unshrunk, coef_scaled_sd):
    """cg_sampler.py:123-138."""
    s = np.ones(len(prior_prec_sqrt))
    s[n_unshrunk:] = prior_prec_sqrt[n_unshrunk:] ** -1
    if n_unshrunk > 0:
        s[:n_unshrunk] = 2. * coef_scaled_sd[:n_unshrunk]
    return s


def cg_sample(design, obs_prec, prior_prec_sqrt, z, x0, s, maxiter, atol, eps1, eps2):
    """cg_sampler.py:55-93 with the Gaussian noise supplied by the caller."""
    D = (s * prior_prec_sqrt) ** 2

    def A(x):
        return D * x + s * design.Tdot(obs_prec * design.dot(s * x))

    v = design.Tdot(obs_prec ** (1 / 2) * eps1) + prior_prec_sqrt * eps2
    b = s * (z + v)
    bnorm = np.linalg.norm(b)
    rtol = atol / bnorm if bnorm > 0 else 0.
    atol_eff = max(0., rtol * bnorm)
    x, info, n_iter = cg_loop(A, b, x0 / s, maxiter, atol_eff)
    return s * x, {'n_iter': n_iter, 'converged': info == 0, 'b': b}


def fisher_full(design, weight):
    """compute_fisher_info(weight, diag_only=False): dense_matrix.py:54-58 on the materialised [1, X - c] image, which
    is also what the sparse class's algebra (sparse_matrix.py:131-162) evaluates to."""
    A = design.toarray()
    return A.T.dot(weight[:, np.newaxis] * A)


def cholesky_sample(design, obs_prec, prior_prec_sqrt, z, gaussian_vec):
    """direct_gaussian_sampler.py:4-44 with the standard normal vector supplied by the caller."""
    import scipy.linalg
    G = fisher_full(design, obs_prec)
    diag = prior_prec_sqrt ** 2 + np.diag(G)
    J = 1 / np.sqrt(diag)
    Prec = J[:, np.newaxis] * G * J[np.newaxis, :]
    Prec += np.diag((J * prior_prec_sqrt) ** 2)
    U = scipy.linalg.cholesky(Prec, lower=False)
    mean = scipy.linalg.cho_solve((U, False), J * z)
    return J * (mean + scipy.linalg.solve_triangular(U, gaussian_vec, lower=False))


def exact_gaussian_mean(design, obs_prec, prior_prec_sqrt, rhs):
    """Dense solve of (X' Omega X + diag(pps^2)) beta = rhs, for small p."""
    A = design.toarray()
    Phi = A.T @ (obs_prec[:, None] * A) + np.diag(prior_prec_sqrt ** 2)
    return np.linalg.solve(Phi, rhs)


# ---- running summaries ---------------------------------------------------------------------------
class SummarizerOracle:
    """reg_coef_posterior_summarizer.py:68-124 + :12-41."""

    def __init__(self, n_coef, n_unshrunk, slab_size, sd_prior_samplesize=5):
        self.k, self.slab = n_unshrunk, slab_size
        self.m, self.q, self.n = np.zeros(n_coef), np.ones(n_coef), 0
        self.n0 = sd_prior_samplesize

    def prior_scale(self, gscale, lscale):
        raw = gscale * lscale
        return raw / np.sqrt(1 + (raw / self.slab) ** 2)

    def update(self, coef, gscale, lscale):
        th = coef.copy()
        th[self.k:] /= self.prior_scale(gscale, lscale)
        w = 1 / (1 + self.n)
        self.m = w * th + (1 - w) * self.m
        self.q = w * th ** 2 + (1 - w) * self.q
        self.n += 1

    def x0(self, gscale, lscale):
        g = self.m.copy()
        g[self.k:] *= self.prior_scale(gscale, lscale)
        return g

    def sd(self):
        if self.n <= 1:
            return np.ones(len(self.m))
        var = self.n / (self.n - 1) * (self.q - self.m ** 2)
        w = (self.n - 1) / (self.n - 1 + self.n0)
        return np.sqrt(w * var + (1 - w) * 1.)


def sample_gaussian_posterior(design, summ, y, obs_prec, gscale, lscale, prior_sd_unshrunk, slab, randn):
    """reg_coef_sampler.py:60-103 ('cg' branch). randn(k) supplies the Gaussian noise."""
    z = design.Tdot(obs_prec * y)
    shrunk = gscale * lscale
    shrunk = shrunk / np.sqrt(1 + (shrunk / slab) ** 2)
    pps = 1 / np.concatenate((prior_sd_unshrunk, shrunk))
    x0 = summ.x0(gscale, lscale)
    s = precond_scale_prior(pps, len(prior_sd_unshrunk), summ.sd())
    eps1 = randn(design.shape[0])
    eps2 = randn(design.shape[1])
    coef, info = cg_sample(design, obs_prec, pps, z, x0, s, 500, 10e-6 * np.sqrt(design.shape[1]), eps1, eps2)
    summ.update(coef, gscale, lscale)
    return coef, info


# ---- the Gibbs loop around it (only what the 'cg' path of the reference's tests touches) ---------
def pg_mean(shape, tilt):
    m = shape.copy() / 2                                    # logistic_model.py:80-87
    nz = np.abs(tilt) > 1e-5
    m[nz] *= 1 / tilt[nz] * (np.exp(tilt[nz]) - 1) / (np.exp(tilt[nz]) + 1)
    return m


def gibbs_cg_oracle(family, outcome, X, n_iter, seed, bridge_exp, sd_for_intercept, slab, init_gscale,
                    init_lscale, pg, ts, sparse_input):
    """Chain of regression_tests/test_gibb.py:26-60 ('cg' combos): init given by global+local scale,
    mode search SKIPPED is not possible there (coef not in init) -- so this restates search_mode too, via
    scipy L-BFGS-B exactly as reg_coef_sampler.py:281-392 sets it up.
    pg / ts: objects with rand_polyagamma(shape, tilt) / sample(char_exp, tilt) (the ports or the reference's)."""
    import scipy.optimize
    design = DesignOracle(X, center_predictor=True, add_intercept=True)
    n, P = design.shape
    np.random.seed(seed)                                    # random.py:17-22
    pg_seed = np.random.randint(1, 1 + np.iinfo(np.int32).max)
    ts_seed = np.random.randint(1, 1 + np.iinfo(np.int32).max)
    pg, ts = pg(pg_seed), ts(ts_seed)
    prior_sd_unshrunk = np.array([sd_for_intercept])
    k = 1
    unit = math.gamma(2 / bridge_exp) / math.gamma(1 / bridge_exp)
    gscale, lscale = init_gscale / unit, init_lscale * unit  # prior.py:128-139 (to 'raw')
    summ = SummarizerOracle(P, k, slab)
    if family == 'logit':
        n_success, n_trial = [np.asarray(a, dtype=float) for a in outcome]
    else:
        y = np.asarray(outcome, dtype=float)

    coef = np.zeros(P)
    if family == 'logit':
        ph = n_success.mean() / n_trial.mean()
        coef[0] = np.log(ph / (1 - ph))
        obs_prec = pg_mean(n_trial, design.dot(coef))
    else:
        coef[0] = y.mean()
        obs_prec = np.mean((y - design.dot(coef)) ** 2) ** -1

    # search_mode: L-BFGS-B in preconditioned coordinates
    shrunk = gscale * lscale
    shrunk = shrunk / np.sqrt(1 + (shrunk / slab) ** 2)
    scale = np.concatenate(([1.], shrunk))
    pprec = np.concatenate(((prior_sd_unshrunk / scale[:k]) ** -2, np.ones(P - k)))

    def loglik_grad(beta, loglik_only):
        eta = design.dot(beta)
        if family == 'logit':
            ll = np.sum(n_success * eta - n_trial * np.logaddexp(0, eta))
            g = None if loglik_only else design.Tdot(n_success - n_trial / (1 + np.exp(-eta)))
        else:
            ll = len(y) * math.log(obs_prec) / 2 - obs_prec * np.sum((y - eta) ** 2) / 2
            g = None if loglik_only else obs_prec * design.Tdot(y - eta)
        return ll, g

    def f(th, loglik_only=False):
        ll, g = loglik_grad(th * scale, loglik_only)
        ll += np.sum(-pprec * th ** 2) / 2
        if g is not None:
            g = scale * g - pprec * th
        return ll, g

    res = scipy.optimize.minimize(lambda t: -f(t, True)[0], coef / scale, jac=lambda t: -f(t)[1],
                                  method='L-BFGS-B',
                                  options={'maxiter': 250, 'gtol': 10 ** -6 / np.sqrt(P), 'maxcor': 200})
    coef = scale * res.x

    def draw_obs_prec(coef):
        if family == 'logit':
            return pg.rand_polyagamma(n_trial.astype(np.intc), design.dot(coef))
        resid = y - design.dot(coef)
        return 1 / (np.sum(resid ** 2) / 2 / np.random.gamma(n / 2, 1))

    def draw_lscale(gscale, coef):
        ls = np.sqrt(.5 / ts.sample(bridge_exp / 2, (coef[k:] / gscale) ** 2))
        return ls

    obs_prec = draw_obs_prec(coef)
    lscale = draw_lscale(gscale, coef)
    lower_bd = .001 / unit
    coefs = np.zeros((P, n_iter))
    n_cg = np.zeros(n_iter, dtype=int)
    for it in range(n_iter):
        if family == 'logit':
            yg, om = (n_success - n_trial / 2) / obs_prec, obs_prec
        else:
            yg, om = y, obs_prec * np.ones(n)
        coef, info = sample_gaussian_posterior(design, summ, yg, om, gscale, lscale, prior_sd_unshrunk, slab,
                                               np.random.randn)
        obs_prec = draw_obs_prec(coef)
        shape = (P - k) / bridge_exp
        rate = np.sum(np.abs(coef[k:]) ** bridge_exp)
        phi = np.random.gamma(shape, scale=1 / rate)
        gscale = max(1 / phi ** (1 / bridge_exp), lower_bd)
        lscale = draw_lscale(gscale, coef)
        coefs[:, it] = coef
        n_cg[it] = info['n_iter']
    return coefs, n_cg
