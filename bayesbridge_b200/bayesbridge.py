"""Gibbs sampler for Bayesian bridge regression (reference: bayesbridge.py), with the coefficient
update (prior-preconditioned CG) and the Polya-Gamma update running in libbbgpu on a B200.

Per iteration the device keeps X, the outcome and the observation precisions omega resident; the host
sees the P-length coefficient vector, the scale parameters and a few scalars."""
import ctypes
import math
import time
from warnings import warn

import numpy as np

from . import _lib
from .random import BasicRandom, DevicePolyaGamma
from .reg_coef_sampler import SparseRegressionCoefficientSampler
from .model import LogisticModel
from .prior import RegressionCoefPrior
from .gibbs_util import MarkovChainManager, SamplerOptions


class BayesBridge():
    """Gibbs sampler for Bayesian bridge sparse regression (linear and logistic likelihoods)."""

    def __init__(self, model, prior=RegressionCoefPrior()):
        self.n_obs = model.n_obs
        self.n_pred = model.n_pred
        self.n_unshrunk = prior.n_fixed
        self.prior_sd_for_unshrunk = prior.sd_for_fixed.copy()
        if model.intercept_added:
            self.n_unshrunk += 1
            self.prior_sd_for_unshrunk = np.concatenate(([prior.sd_for_intercept], self.prior_sd_for_unshrunk))
        self.model = model
        self.prior = prior
        self.rg = BasicRandom(ctx=model.design.ctx)
        self.manager = MarkovChainManager(self.n_obs, self.n_pred, self.n_unshrunk, model.name)
        self._push_outcome()

    # ---- device residency ---------------------------------------------------------------------
    def _push_outcome(self):
        """Make the outcome resident next to X so that X'(omega*y), the PG tilt and the log-likelihood
        never need an n-length host round trip."""
        lib, mat = _lib.load(), self.model.design._mat
        if self.model.name == 'logit':
            _lib.check(lib.bb_set_outcome(mat, _lib.dptr(_lib.as_f64(self.model.n_trial)),
                                          _lib.dptr(_lib.as_f64(self.model.n_success))))
        else:
            _lib.check(lib.bb_set_outcome(mat, None, _lib.dptr(_lib.as_f64(self.model.y))))
        # the outcome, the cached X'kappa and the P-side state live on the DESIGN's device handle, and a design may be
        # shared by several models / bridges (several outcomes on one X): remember whose outcome is resident
        self.model.design._outcome_owner = id(self)
        self.model.design._outcome_model = id(self.model)

    def _ensure_outcome(self):
        """Re-push this bridge's outcome if another bridge bound to the same design pushed its own since
        (bb_set_outcome also invalidates the cached X'kappa)."""
        if getattr(self.model.design, '_outcome_owner', None) != id(self):
            self._push_outcome()

    def _fetch_obs_prec(self):
        out = np.empty(self.n_obs)
        _lib.check(_lib.load().bb_get_obs_prec(self.model.design._mat, _lib.dptr(out)))
        return out

    @property
    def _pg_on_device(self):
        return isinstance(self.rg.pg, DevicePolyaGamma)

    # ---- device-resident P-side state (lambda, running summaries): SURVEY section 8f-2 -----------------------------
    def _resident_mode(self, options):
        """The P-side Gibbs state can stay on the device when every sampler involved is the device one."""
        import os
        from .random import DeviceTiltedStable
        return (options.noise == 'device' and options.coef_sampler_type == 'cg'
                and self._pg_on_device and isinstance(self.rg.ts, DeviceTiltedStable)
                and self.prior.bridge_exp != 2 and (self.n_pred - self.n_unshrunk) > 0
                and os.environ.get('BB_RESIDENT_STATE', '1') != '0')

    def _state_push(self, lscale):
        lib, mat = _lib.load(), self.model.design._mat
        summ = self.reg_coef_sampler.regcoef_summarizer.coef_scaled_summarizer
        _lib.check(lib.bb_state_init(mat, int(self.n_unshrunk), _lib.dptr(_lib.as_f64(self.prior_sd_for_unshrunk)),
                                     float(self.prior.slab_size)))
        _lib.check(lib.bb_state_set(mat, _lib.dptr(_lib.as_f64(lscale)), _lib.dptr(_lib.as_f64(summ.stats['mean'])),
                                    _lib.dptr(_lib.as_f64(summ.stats['square'])), int(summ.n_averaged)))

    def _state_pull(self, want_lscale=True):
        """Copy lambda and the summaries back into the host objects (end of a run, or when lambda is saved)."""
        lib, mat = _lib.load(), self.model.design._mat
        summ = self.reg_coef_sampler.regcoef_summarizer.coef_scaled_summarizer
        lscale = np.empty(self.n_pred - self.n_unshrunk) if want_lscale else None
        mean, square = np.empty(self.n_pred), np.empty(self.n_pred)
        n_avg = ctypes.c_int64()
        _lib.check(lib.bb_state_get(mat, _lib.dptr(lscale), _lib.dptr(mean), _lib.dptr(square), ctypes.byref(n_avg)))
        summ.stats['mean'], summ.stats['square'], summ.n_averaged = mean, square, int(n_avg.value)
        return lscale

    def _resident_iteration(self, obs_prec, gscale, options, need_lscale):
        """One Gibbs iteration with lambda / summaries on the device. Returns coef, obs_prec, gscale, lscale, logp, info."""
        lib, design = _lib.load(), self.model.design
        mat, bridge_exp = design._mat, self.prior.bridge_exp
        P, k = self.n_pred, self.n_unshrunk
        if self.model.name == 'linear':
            _lib.check(lib.bb_set_obs_prec_scalar(mat, float(obs_prec)))
            omega = None
        else:
            omega = None if obs_prec is _RESIDENT else _lib.as_f64(obs_prec)
        coef, sums = np.empty(P), np.empty(4)
        n_iter, info = ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.bb_cg_sample_resident(
            mat, _lib.dptr(omega), float(gscale), float(bridge_exp), 10e-6 * np.sqrt(P), 500,
            self.rg.cg.seed, self.rg.cg._next_offset(), _lib.dptr(coef), ctypes.byref(n_iter), ctypes.byref(info),
            _lib.dptr(sums)))
        design.dot_count += n_iter.value + 1
        design.Tdot_count += n_iter.value + 2
        if info.value != 0:
            warn("The conjugate gradient algorithm did not achieve the requested tolerance level. You may "
                 "increase the maxiter or use the dense linear algebra instead.")
        obs_prec = self.update_obs_precision(coef, coef_is_resident=True)
        # tau | beta from the device-side sums (bayesbridge.py:412-448)
        abs_pow_sum, n_nonzero, slab_sq_sum, unshrunk_sq_sum = sums
        lower_bd = .001 / self.prior.compute_power_exp_ave_magnitude(bridge_exp)
        if options.gscale_update == 'sample':
            if n_nonzero == 0:
                gscale = 0
            else:
                hyper = self.prior.param['gscale_neg_power']
                shape = hyper['shape'] + (P - k) / bridge_exp
                rate = hyper['rate'] + abs_pow_sum
                phi = self.rg.np_random.gamma(shape, scale=1 / rate)
                gscale = 1 / phi ** (1 / bridge_exp)
        elif options.gscale_update == 'optimize':
            gscale = ((P - k) / bridge_exp / abs_pow_sum) ** - (1 / bridge_exp)
        if (options.gscale_update is not None) and gscale < lower_bd:
            gscale = lower_bd
            warn("The global shrinkage parameter update returned an unreasonably "
                 "small value. Returning a specified lower bound value instead.")
        # lambda | tau, beta on the device
        counts = (ctypes.c_int * 3)()
        lscale = np.empty(P - k) if need_lscale else None
        _lib.check(lib.bb_local_scale_resident(mat, float(gscale), bridge_exp / 2, self.rg.ts.seed,
                                               self.rg.ts._next_offset(), counts, _lib.dptr(lscale)))
        if counts[0] > 0:
            raise ValueError('Tilting parameter must be positive.')
        if counts[1] > 0:
            warn("Local scale parameter under-flowed. Replacing with a small number.")
        elif counts[2] > 0:
            warn("Local scale parameter over-flowed. Replacing with a large number.")
        # log posterior (bayesbridge.py:480-511) from the same sums; sum|b/tau|^a = sum|b|^a / tau^a
        if self.model.name == 'logit':
            loglik = self._loglik_cache[1]
        else:
            # linear_model.py:22-29 with the residual sum of squares update_obs_precision just reduced on the device
            loglik = self.model.n_obs_global * math.log(obs_prec) / 2 - obs_prec * self._last_rss / 2
        if not np.isinf(self.prior.slab_size):
            loglik += - .5 * slab_sq_sum
        prior_logp = - (P - k) * math.log(gscale) - abs_pow_sum / gscale ** bridge_exp
        prior_logp += - 1 / 2 * unshrunk_sq_sum
        prior_logp += - np.sum(np.log(self.prior_sd_for_unshrunk[self.prior_sd_for_unshrunk < float('inf')]))
        hyper = self.prior.param['gscale_neg_power']
        prior_logp += (hyper['shape'] - 1.) * math.log(gscale) - hyper['rate'] * gscale
        return coef, obs_prec, gscale, lscale, loglik + prior_logp, {'n_cg_iter': n_iter.value}

    # ---- public API ---------------------------------------------------------------------------
    def gibbs_resume(self, prev_mcmc_info, n_add_iter, n_status_update=0, merge=False, prev_samples=None):
        """Continue a chain from the state stored in `prev_mcmc_info` (reference: bayesbridge.py:43-107)."""
        if merge and prev_samples is None:
            raise ValueError(
                "To merge the outputs from previous and new MCMC runs, you "
                "have to supply the optional argument `prev_samples`.")
        self.rg.set_state(prev_mcmc_info['_random_gen_state'])
        self.reg_coef_sampler = SparseRegressionCoefficientSampler(
            self.n_pred, self.prior_sd_for_unshrunk, prev_mcmc_info['coef_sampler_type'],
            prev_mcmc_info['options']['hmc_curvature_est_stabilized'], self.prior.slab_size)
        self.reg_coef_sampler.set_internal_state(prev_mcmc_info['_reg_coef_sampler_state'])
        init = dict(prev_mcmc_info['_markov_chain_state'])
        if '_markov_chain_state_raw_scales' in prev_mcmc_info:
            init['_raw_scales'] = prev_mcmc_info['_markov_chain_state_raw_scales']
        new_samples, new_mcmc_info = self.gibbs(
            n_add_iter, 0, prev_mcmc_info['thin'], init=init,
            params_to_save=prev_mcmc_info['saved_params'], n_status_update=n_status_update,
            options=prev_mcmc_info['options'], _add_iter_mode=True)
        if merge:
            new_samples, new_mcmc_info = self.manager.merge_outputs(
                prev_samples, prev_mcmc_info, new_samples, new_mcmc_info)
        return new_samples, new_mcmc_info

    def gibbs(self, n_iter, n_burnin=0, thin=1, seed=None, init={'global_scale': 0.1},
              params_to_save=('coef', 'global_scale', 'logp'), coef_sampler_type=None,
              n_status_update=0, options=None, _add_iter_mode=False):
        """Generate posterior samples. Arguments and return values as in the reference's
        `BayesBridge.gibbs` (bayesbridge.py:109-277); `options` additionally accepts
        {'noise': 'device' | 'host'} (see SamplerOptions).

        Returns (samples, mcmc_info): samples[param][..., i] is the i-th saved draw."""
        if not isinstance(options, SamplerOptions):
            options = SamplerOptions.pick_default_and_create(
                coef_sampler_type, options, self.model.name, self.model.design)
        self._ensure_outcome()
        if not _add_iter_mode:
            ctx = self.model.design.ctx
            if seed is None and ctx.nranks > 1:
                # the P-side state is replicated: every rank must draw the same tau, so they must share a seed
                mine = float(np.random.SeedSequence().generate_state(1)[0]) if ctx.rank == 0 else 0.0
                seed = int(ctx.allreduce_host(np.array([mine]))[0])
            self.rg.set_seed(seed)
            self.reg_coef_sampler = SparseRegressionCoefficientSampler(
                self.n_pred, self.prior_sd_for_unshrunk, options.coef_sampler_type,
                options.curvature_est_stabilized, self.prior.slab_size)
            self.reg_coef_sampler.init_optimizer = options.init_optimizer
        if params_to_save == 'all':
            params_to_save = ('coef', 'local_scale', 'global_scale', 'logp', 'obs_prec')
        n_status_update = min(n_iter, n_status_update)
        start_time = time.time()
        self.manager.stamp_time(start_time)

        coef, obs_prec, lscale, gscale, init, initial_optim_info = \
            self.initialize_chain(init, self.prior.bridge_exp)
        init_runtime = time.time() - start_time          # chain initialisation (mode search) split out of 'runtime'
        self._loglik_cache = None

        samples, sampling_info = {}, {}
        self.manager.pre_allocate(
            samples, sampling_info, n_iter - n_burnin, thin, params_to_save, options.coef_sampler_type)

        resident = self._resident_mode(options) and n_iter > 0
        if resident:
            self._state_push(lscale)
        save_lscale = 'local_scale' in params_to_save
        for mcmc_iter in range(1, n_iter + 1):
            if resident:
                coef, obs_prec, gscale, lscale_new, logp, info = self._resident_iteration(
                    obs_prec, gscale, options,
                    need_lscale=save_lscale and self.manager._slot(mcmc_iter, n_burnin, thin) is not None)
                if lscale_new is not None:
                    lscale = lscale_new
                self.manager.store_current_state(
                    samples, mcmc_iter, n_burnin, thin, coef, lscale, gscale,
                    self._host_obs_prec_getter(obs_prec), logp, params_to_save)
                self.manager.store_sampling_info(
                    sampling_info, info, mcmc_iter, n_burnin, thin, options.coef_sampler_type)
                self.manager.print_status(n_status_update, mcmc_iter, n_iter)
                continue
            coef, info = self.update_regress_coef(
                coef, obs_prec, gscale, lscale, options.coef_sampler_type, noise=options.noise)
            obs_prec = self.update_obs_precision(coef)
            # tau | beta first, then lambda | tau, beta: the order matters
            gscale = self.update_global_scale(
                gscale, coef[self.n_unshrunk:], self.prior.bridge_exp, method=options.gscale_update)
            lscale = self.update_local_scale(gscale, coef[self.n_unshrunk:], self.prior.bridge_exp)
            logp = self.compute_posterior_logprob(coef, gscale, obs_prec, self.prior.bridge_exp)
            self.manager.store_current_state(
                samples, mcmc_iter, n_burnin, thin, coef, lscale, gscale,
                self._host_obs_prec_getter(obs_prec), logp, params_to_save)
            self.manager.store_sampling_info(
                sampling_info, info, mcmc_iter, n_burnin, thin, options.coef_sampler_type)
            self.manager.print_status(n_status_update, mcmc_iter, n_iter)

        if resident:
            lscale = self._state_pull(want_lscale=True)      # lambda and the summaries back into the host objects
        runtime = time.time() - start_time
        ctx = self.model.design.ctx
        if ctx.nranks > 1:
            ready, err = ctx.p2p_status()
            if err:
                raise RuntimeError("peer-memory exchange timed out waiting for another rank; results are invalid")
        raw_scales = {'global_scale': gscale, 'local_scale': np.array(lscale, copy=True)}

        if self.prior._gscale_paramet == 'coef_magnitude':
            gscale, lscale = self.prior.adjust_scale(gscale, lscale, to='coef_magnitude')
            self.prior.adjust_scale(
                samples.get('global_scale', 0.), samples.get('local_scale', 0.), to='coef_magnitude')

        obs_prec_host = self._host_obs_prec_getter(obs_prec)
        obs_prec_host = obs_prec_host() if callable(obs_prec_host) else obs_prec_host
        mcmc_info = {
            'init': init, 'n_iter': n_iter, 'n_burnin': n_burnin, 'thin': thin, 'seed': seed,
            'n_coef_wo_shrinkage': self.n_unshrunk,
            'prior_sd_for_unshrunk': self.prior_sd_for_unshrunk,
            'bridge_exponent': self.prior.bridge_exp,
            'coef_sampler_type': options.coef_sampler_type,
            'saved_params': params_to_save,
            'runtime': runtime,
            'init_runtime': init_runtime,
            'options': options.get_info(),
            '_init_optim_info': initial_optim_info,
            '_reg_coef_sampling_info': sampling_info,
            '_markov_chain_state': self.manager.pack_parameters(coef, obs_prec_host, lscale, gscale),
            # the sampler's own (raw) parametrisation: tau*u/u is not a floating-point identity, and a one-ulp
            # change is enough to make a resumed chain differ from an uninterrupted one
            '_markov_chain_state_raw_scales': raw_scales,
            '_random_gen_state': self.rg.get_state(),
            '_reg_coef_sampler_state': self.reg_coef_sampler.get_internal_state(),
        }
        return samples, mcmc_info

    # ---- initial state --------------------------------------------------------------------------
    def initialize_chain(self, init, bridge_exp):
        """User-specified state where given, heuristics / conditional optimisation elsewhere
        (reference: bayesbridge.py:279-353)."""
        for key in init:
            if key not in ('coef', 'local_scale', 'global_scale', 'obs_prec', 'logp', '_raw_scales'):
                warn("'{:s}' is not a valid parameter name and will be ignored.".format(key))
        n_shrunk = self.n_pred - self.n_unshrunk
        have_coef = 'coef' in init
        if have_coef:
            coef = np.array(init['coef'], dtype=np.float64)
            if len(coef) != self.n_pred:
                raise ValueError('Invalid initial length of regression coefficient.')
        else:
            coef = np.zeros(self.n_pred)
            if self.model.intercept_added:
                coef[0] = self.model.calc_intercept_mle()

        obs_prec = self.initialize_obs_precision(init, coef)

        if have_coef and 'global_scale' not in init:
            gscale = self.update_global_scale(None, coef[self.n_unshrunk:], bridge_exp, method='optimize')
            lscale = self.update_local_scale(gscale, coef[self.n_unshrunk:], bridge_exp)
        else:
            if 'global_scale' not in init:
                raise ValueError("Initial global scale must be specified when coefficients aren't specified.")
            if self.prior._gscale_paramet == 'raw':
                warn("Using the raw global scale parametrization; make sure that "
                     "the specified initial value is scaled accordingly.")
            gscale = init['global_scale']
            if 'local_scale' in init:
                lscale = np.array(init['local_scale'], dtype=np.float64)
                if len(lscale) != n_shrunk:
                    raise ValueError('Invalid initial length of local scale parameter')
            else:
                lscale = np.ones(n_shrunk)

        if '_raw_scales' in init:
            gscale = init['_raw_scales']['global_scale']
            lscale = np.array(init['_raw_scales']['local_scale'], dtype=np.float64)
        elif self.prior._gscale_paramet == 'coef_magnitude':
            # the sampler itself works in the raw parametrisation
            gscale, lscale = self.prior.adjust_scale(gscale, lscale, to='raw')

        optim_info = None
        if not have_coef:
            obs_prec_host = self._host_obs_prec_getter(obs_prec)
            obs_prec_host = obs_prec_host() if callable(obs_prec_host) else obs_prec_host
            coef, info = self.reg_coef_sampler.search_mode(coef, lscale, gscale, obs_prec_host, self.model)
            obs_prec = self.update_obs_precision(coef)
            lscale = self.update_local_scale(gscale, coef[self.n_unshrunk:], bridge_exp)
            optim_info = {key: info[key] for key in ['is_success', 'n_design_matvec', 'n_iter']}

        obs_prec_host = self._host_obs_prec_getter(obs_prec)
        init = {'coef': coef, 'obs_prec': obs_prec_host() if callable(obs_prec_host) else obs_prec_host,
                'local_scale': lscale, 'global_scale': gscale}
        return coef, obs_prec, lscale, gscale, init, optim_info

    def initialize_obs_precision(self, init, coef):
        if 'obs_prec' in init and init['obs_prec'] is not None:
            obs_prec = init['obs_prec']
            if self.model.name == 'logit':
                obs_prec = np.array(obs_prec, dtype=np.float64, order='C')
                if len(obs_prec) != self.n_obs:
                    raise ValueError('An invalid initial state.')
        elif self.model.name == 'linear':
            rss = self._linear_rss(coef)
            obs_prec = (rss / self.model.n_obs_global) ** -1
        else:
            obs_prec = LogisticModel.compute_polya_gamma_mean(
                self.model.n_trial, self.model.design.dot(coef))
        return obs_prec

    # ---- conditional updates ----------------------------------------------------------------------
    def update_regress_coef(self, coef, obs_prec, gscale, lscale, sampling_method, noise='device'):
        """beta | omega, tau, lambda by the CG sampler or the direct (Cholesky) draw (reference: bayesbridge.py:372-395)."""
        if sampling_method not in ('cg', 'cholesky'):
            raise NotImplementedError()
        lib, mat = _lib.load(), self.model.design._mat
        if sampling_method == 'cholesky':
            design = self.model.design
            if self.model.name == 'linear':
                _lib.check(lib.bb_set_obs_prec_scalar(mat, float(obs_prec)))
                z = design.Tdot(float(obs_prec) * _lib.as_f64(self.model.y))
                omega = None
            else:
                z = design.Tdot(_lib.as_f64(self.model.n_success - self.model.n_trial / 2))   # omega * (kappa / omega)
                omega = None if obs_prec is _RESIDENT else _lib.as_f64(obs_prec)
            return self.reg_coef_sampler.sample_gaussian_posterior(None, design, omega, gscale, lscale, 'cholesky', z=z)
        philox = (self.rg.cg.seed, self.rg.cg._next_offset()) if noise == 'device' else None
        if self.model.name == 'linear':
            # omega = sigma^-2 * 1_n: the device keeps it as a scalar; z = omega X'y is formed there
            _lib.check(lib.bb_set_obs_prec_scalar(mat, float(obs_prec)))
            return self.reg_coef_sampler.sample_gaussian_posterior(
                None, self.model.design, None, gscale, lscale, sampling_method, noise=noise, philox=philox)
        if obs_prec is _RESIDENT:
            # omega already on the device (left there by the fused PG update); z = X' kappa cached there
            return self.reg_coef_sampler.sample_gaussian_posterior(
                None, self.model.design, None, gscale, lscale, sampling_method, noise=noise, philox=philox)
        # host omega (initial / resumed state, or a host-side PG sampler was plugged in): omega is uploaded.
        if noise == 'host':
            # parity mode: the reference's arithmetic, z = X'(omega * (kappa / omega))  (bayesbridge.py:380-383)
            y_gaussian = (self.model.n_success - self.model.n_trial / 2) / obs_prec
        else:
            # omega * y_gaussian == kappa, so z = X'kappa is formed (once) on the device; using the same z
            # as the resident path keeps a resumed chain bit-identical to an uninterrupted one
            y_gaussian = None
        return self.reg_coef_sampler.sample_gaussian_posterior(
            y_gaussian, self.model.design, _lib.as_f64(obs_prec), gscale, lscale, sampling_method,
            noise=noise, philox=philox)

    def update_obs_precision(self, coef, coef_is_resident=False):
        """omega | beta (reference: bayesbridge.py:397-410). coef_is_resident: the device still holds these very
        coefficients from the CG draw, so they are not uploaded again."""
        if self.model.name == 'linear':
            scale = self._linear_rss(coef, coef_is_resident) / 2
            obs_var = scale / self.rg.np_random.gamma(self.model.n_obs_global / 2, 1)
            return 1 / obs_var
        if self._pg_on_device:
            # fused on the device: eta = X beta, omega ~ PG(n_trial, eta), log-likelihood; omega stays there
            loglik = ctypes.c_double()
            _lib.check(_lib.load().bb_pg_from_coef(
                self.model.design._mat, None if coef_is_resident else _lib.dptr(_lib.as_f64(coef)),
                self.rg.pg.seed, self.rg.pg._next_offset(), None, ctypes.byref(loglik)))
            self.model.design.dot_count += 1
            self._loglik_cache = (coef, loglik.value)
            return _RESIDENT
        return self.rg.polya_gamma(self.model.n_trial.astype(np.intc), self.model.design.dot(coef),
                                   index_offset=getattr(self.model.design, 'row_offset', 0))

    def _linear_rss(self, coef, coef_is_resident=False):
        rss = ctypes.c_double()
        _lib.check(_lib.load().bb_linear_rss(
            self.model.design._mat, None if coef_is_resident else _lib.dptr(_lib.as_f64(coef)), ctypes.byref(rss)))
        self.model.design.dot_count += 1
        self._last_rss = rss.value
        return rss.value

    def _host_obs_prec_getter(self, obs_prec):
        return self._fetch_obs_prec if obs_prec is _RESIDENT else obs_prec

    def update_global_scale(self, gscale, coef_under_shrinkage, bridge_exp,
                            coef_expected_magnitude_lower_bd=.001, method='sample'):
        """tau | beta: conjugate Gamma update of phi = tau^-alpha (reference: bayesbridge.py:412-448)."""
        if coef_under_shrinkage.size == 0:
            return 1.
        lower_bd = coef_expected_magnitude_lower_bd / self.prior.compute_power_exp_ave_magnitude(bridge_exp)
        if method == 'optimize':
            gscale = self.monte_carlo_em_global_scale(coef_under_shrinkage, bridge_exp)
        elif method == 'sample':
            if np.count_nonzero(coef_under_shrinkage) == 0:
                gscale = 0
            else:
                hyper = self.prior.param['gscale_neg_power']
                shape = hyper['shape'] + coef_under_shrinkage.size / bridge_exp
                rate = hyper['rate'] + np.sum(np.abs(coef_under_shrinkage) ** bridge_exp)
                phi = self.rg.np_random.gamma(shape, scale=1 / rate)
                gscale = 1 / phi ** (1 / bridge_exp)
        if (method is not None) and gscale < lower_bd:
            gscale = lower_bd
            warn("The global shrinkage parameter update returned an unreasonably "
                 "small value. Returning a specified lower bound value instead.")
        return gscale

    def monte_carlo_em_global_scale(self, coef_under_shrinkage, bridge_exp):
        phi = len(coef_under_shrinkage) / bridge_exp / np.sum(np.abs(coef_under_shrinkage) ** bridge_exp)
        return phi ** - (1 / bridge_exp)

    def update_local_scale(self, gscale, coef_under_shrinkage, bridge_exp):
        """lambda | tau, beta through exponentially tilted stable draws (reference: bayesbridge.py:458-478)."""
        if bridge_exp == 2:
            return .5 * np.ones(coef_under_shrinkage.size)
        lscale_sq = .5 / self.rg.tilted_stable(bridge_exp / 2, (coef_under_shrinkage / gscale) ** 2)
        lscale = np.sqrt(lscale_sq)
        if np.any(lscale == 0):
            warn("Local scale parameter under-flowed. Replacing with a small number.")
            lscale[lscale == 0] = 10e-16
        elif np.any(np.isinf(lscale)):
            warn("Local scale parameter over-flowed. Replacing with a large number.")
            lscale[np.isinf(lscale)] = 2.0 / gscale
        return lscale

    def compute_posterior_logprob(self, coef, gscale, obs_prec, bridge_exp):
        """Log posterior density up to a constant (reference: bayesbridge.py:480-511)."""
        cache = getattr(self, '_loglik_cache', None)
        if self.model.name == 'logit' and cache is not None and cache[0] is coef:
            loglik = cache[1]          # computed by the fused PG kernel on the same coef
        elif self.model.name == 'linear':
            loglik, _ = self.model.compute_loglik_and_gradient(coef, obs_prec, loglik_only=True)
        else:
            loglik, _ = self.model.compute_loglik_and_gradient(coef, loglik_only=True)
        if not np.isinf(self.prior.slab_size):
            loglik += - .5 * np.sum((coef / self.prior.slab_size) ** 2)

        n_shrunk = len(coef) - self.n_unshrunk
        prior_logp = - n_shrunk * math.log(gscale) \
            - np.sum(np.abs(coef[self.n_unshrunk:] / gscale) ** bridge_exp)
        prior_logp += - 1 / 2 * np.sum((coef[:self.n_unshrunk] / self.prior_sd_for_unshrunk) ** 2)
        prior_logp += - np.sum(np.log(
            self.prior_sd_for_unshrunk[self.prior_sd_for_unshrunk < float('inf')]))
        hyper = self.prior.param['gscale_neg_power']
        prior_logp += (hyper['shape'] - 1.) * math.log(gscale) - hyper['rate'] * gscale
        return loglik + prior_logp


class _Resident:
    """Marker: the observation precisions of the current state live on the device."""

    def __repr__(self):
        return '<obs_prec resident on device>'


_RESIDENT = _Resident()
