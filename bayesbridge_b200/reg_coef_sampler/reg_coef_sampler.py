"""Coefficient update of the Gibbs sampler (reference: reg_coef_sampler/reg_coef_sampler.py).

Only the Gaussian-posterior path (linear / logistic likelihood, 'cg' method) is offered: that is
the path this package re-implements on the GPU.  `search_mode` (chain initialisation) runs scipy's
L-BFGS-B on the host and reaches the device through the design matrix' dot / Tdot."""
from warnings import warn

import numpy as np
import scipy.optimize

from .cg_sampler import ConjugateGradientSampler
from .direct_gaussian_sampler import generate_gaussian_with_weight
from .reg_coef_posterior_summarizer import RegressionCoeffficientPosteriorSummarizer


class SparseRegressionCoefficientSampler():

    def __init__(self, n_coef, prior_sd_for_unshrunk, sampling_method,
                 stability_estimate_stabilized=False, regularizing_slab_size=float('inf')):
        if sampling_method not in ('cg', 'cholesky'):
            raise ValueError("Only the 'cg' and 'cholesky' samplers are implemented for device-resident design matrices.")
        self.prior_sd_for_unshrunk = prior_sd_for_unshrunk
        self.n_unshrunk = len(prior_sd_for_unshrunk)
        self.regularizing_slab_size = regularizing_slab_size
        self.regcoef_summarizer = RegressionCoeffficientPosteriorSummarizer(
            n_coef, self.n_unshrunk, regularizing_slab_size)
        self.cg_sampler = ConjugateGradientSampler(self.n_unshrunk)
        self._sampling_info_attributes = ['regcoef_summarizer']
        # 'device': the L-BFGS mode search runs inside libbbgpu (bb_mode_search); 'scipy': scipy's L-BFGS-B on the host
        # with the likelihood evaluated on the device (the reference's optimiser, reg_coef_sampler.py:296-305)
        self.init_optimizer = 'device'

    def get_internal_state(self):
        return {a: getattr(self, a) for a in self._sampling_info_attributes if hasattr(self, a)}

    def set_internal_state(self, state):
        for a in self._sampling_info_attributes:
            if hasattr(self, a) and a in state:
                setattr(self, a, state[a])

    def compute_prior_shrunk_scale(self, gscale, lscale):
        """tau*lambda damped by the slab (reg_coef_sampler.py:194-201)."""
        scale = gscale * lscale
        if not np.isinf(self.regularizing_slab_size):       # without a slab the damping factor is exactly 1
            scale /= np.sqrt(1 + (scale / self.regularizing_slab_size) ** 2)
        return scale

    def sample_gaussian_posterior(self, y, design, obs_prec, gscale, lscale, method='cg',
                                  noise='host', philox=None, z=None):
        """
        beta | omega, tau, lambda ~ N(Phi^{-1} X' Omega y, Phi^{-1}) (reg_coef_sampler.py:60-103).

        y : array, or None when the outcome is resident on the device (bb_set_outcome): z = X'(Omega y) is
            then formed there (for the logit model X'kappa, computed once and cached).
        obs_prec : array, or None to use the precision vector resident on the device.
        """
        if method not in ('cg', 'cholesky'):
            raise NotImplementedError("Only method='cg' and 'cholesky' are available on the device.")
        if z is None and y is not None:
            z = design.Tdot(obs_prec * y)
        if method == 'cholesky':
            # reg_coef_sampler.py:74-85: the direct draw needs no initial guess and leaves the summaries alone
            prior_sd = np.concatenate((np.asarray(self.prior_sd_for_unshrunk, dtype=np.float64),
                                       self.compute_prior_shrunk_scale(gscale, lscale)))
            coef = generate_gaussian_with_weight(design, obs_prec, 1 / prior_sd, z)
            return coef, {}
        k = self.n_unshrunk
        prior_prec_sqrt = np.empty(k + len(lscale))
        prior_prec_sqrt[:k] = 1 / np.asarray(self.prior_sd_for_unshrunk, dtype=np.float64)
        np.divide(1., self.regcoef_summarizer.compute_prior_scale(gscale, lscale), out=prior_prec_sqrt[k:])
        x0 = self.regcoef_summarizer.extrapolate_coef_condmean(gscale, lscale)
        scaled_sd = self.regcoef_summarizer.estimate_coef_precond_scale_sd()
        coef, cg_info = self.cg_sampler.sample(
            design, obs_prec, prior_prec_sqrt, z,
            coef_cg_init=x0, precond_by='prior', coef_scaled_sd=scaled_sd,
            maxiter=500, atol=10e-6 * np.sqrt(design.shape[1]),
            noise=noise, philox=philox)
        self.regcoef_summarizer.update(coef, gscale, lscale)
        return coef, {'n_cg_iter': cg_info['n_iter']}

    def _search_mode_on_device(self, coef, scale, prior_prec, obs_prec, model, maxiter, gtol, warn_optim_failure):
        """The same search (same objective, memory, tolerances and stopping rules as the reference's L-BFGS-B call) without
        the host: scipy's own bookkeeping is 1.4 s of a 1.6 s initialisation at P = 1e5 (profiles/r02_chain_init_profile.log)."""
        import ctypes
        from .. import _lib
        design = model.design
        model._ensure_outcome_resident()
        design.reset_matvec_count()
        out = np.empty(coef.size)
        n_iter, n_eval, status = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        prec = float(obs_prec) if model.name == 'linear' else 1.0
        _lib.check(_lib.load().bb_mode_search(
            design._mat, _lib.dptr(_lib.as_f64(coef)), _lib.dptr(_lib.as_f64(scale)), _lib.dptr(_lib.as_f64(prior_prec)),
            prec, int(maxiter), float(gtol), 2.220446049250313e-09, 200,      # ftol = factr * eps, scipy's default
            _lib.dptr(out), ctypes.byref(n_iter), ctypes.byref(n_eval), ctypes.byref(status)))
        design.dot_count += n_eval.value
        design.Tdot_count += n_eval.value
        success = status.value in (0, 1)
        if (not success) and warn_optim_failure:
            warn("The regression coefficient mode could not be located within {:d} optimization "
                 "steps. Proceeding with the current best estimate.".format(n_iter.value))
        info = {
            'is_success': success, 'method': 'L-BFGS (device)', 'n_iter': n_iter.value,
            'n_logp_eval': n_eval.value, 'n_grad_eval': n_eval.value, 'n_hess_eval': 0,
            'n_design_matvec': design.n_matvec,
        }
        return out, info

    # ---- chain initialisation ---------------------------------------------------------------
    def compute_preconditioning_scale(self, gscale, lscale, post_sd, prior_sd_for_unshrunk, target=1.):
        n_coef = len(post_sd)
        k = n_coef - len(lscale)
        scale = np.ones(n_coef)
        scale[k:] = self.compute_prior_shrunk_scale(gscale, lscale)
        if k > 0:
            scale[:k] = target * post_sd[:k]
        prior_prec = np.concatenate(((prior_sd_for_unshrunk / scale[:k]) ** -2, np.ones(len(lscale))))
        return scale, prior_prec

    def search_mode(self, coef, lscale, gscale, obs_prec, model, optim_maxiter=None,
                    warn_optim_failure=False):
        """Conditional posterior mode of the coefficients by L-BFGS-B in prior-preconditioned
        coordinates (reg_coef_sampler.py:281-327)."""
        scale, prior_prec = self.compute_preconditioning_scale(
            gscale, lscale, np.ones(coef.size), self.prior_sd_for_unshrunk)
        loglik_args = (obs_prec,) if model.name == 'linear' else ()
        maxiter = 250 if optim_maxiter is None else optim_maxiter
        tol = 10 ** -6 / np.sqrt(len(coef))
        if self.init_optimizer == 'device' and getattr(model.design, '_mat', None) is not None:
            return self._search_mode_on_device(coef, scale, prior_prec, obs_prec, model, maxiter, tol, warn_optim_failure)

        def logp_and_grad(theta, loglik_only=False):
            logp, grad = model.compute_loglik_and_gradient(scale * theta, *loglik_args, loglik_only=loglik_only)
            logp += np.sum(- prior_prec * theta ** 2) / 2
            if grad is not None and np.isfinite(logp):
                grad = scale * grad - prior_prec * theta
            else:
                grad = None
            return logp, grad

        design = model.design
        design.memoize_dot(True)
        design.reset_matvec_count()
        res = scipy.optimize.minimize(
            lambda th: - logp_and_grad(th, loglik_only=True)[0], coef / scale,
            jac=lambda th: - logp_and_grad(th)[1], method='L-BFGS-B',
            options={'maxiter': maxiter, 'gtol': tol, 'maxcor': 200})
        design.memoize_dot(False)
        if (not res.success) and warn_optim_failure:
            warn("The regression coefficient mode could not be located within {:d} optimization "
                 "steps. Proceeding with the current best estimate.".format(res.nit))
        info = {
            'is_success': res.success, 'method': 'L-BFGS-B', 'n_iter': res['nit'],
            'n_logp_eval': res['nfev'], 'n_grad_eval': res.get('njev', 0), 'n_hess_eval': 0,
            'n_design_matvec': design.n_matvec,
        }
        return scale * res.x, info
