"""Sampler options and MCMC output bookkeeping (reference: gibbs_util.py)."""
import math
import time
from warnings import warn

import numpy as np


class SamplerOptions():

    def __init__(self, coef_sampler_type, global_scale_update='sample',
                 hmc_curvature_est_stabilized=False, noise='device', init_optimizer='device'):
        """
        coef_sampler_type : 'cg' (the device path, default) or 'cholesky' (direct draw on the device); 'hmc' is rejected
        global_scale_update : 'sample' | 'optimize' | None
        noise : 'device' -- CG right-hand-side noise from on-device Philox streams (default);
                'host'   -- drawn from np.random exactly like the reference and injected (parity mode)
        init_optimizer : 'device' -- the L-BFGS mode search of the chain initialisation inside libbbgpu (default);
                         'scipy'  -- scipy's L-BFGS-B on the host, as the reference (reg_coef_sampler.py:296-305)
        """
        if coef_sampler_type not in ('cholesky', 'cg', 'hmc'):
            raise ValueError("Unsupported regression coefficient sampler.")
        if noise not in ('device', 'host'):
            raise ValueError("noise must be 'device' or 'host'.")
        if init_optimizer not in ('device', 'scipy'):
            raise ValueError("init_optimizer must be 'device' or 'scipy'.")
        self.init_optimizer = init_optimizer
        self.coef_sampler_type = coef_sampler_type
        self.gscale_update = global_scale_update
        self.curvature_est_stabilized = hmc_curvature_est_stabilized
        self.noise = noise

    def get_info(self):
        return {
            'coef_sampler_type': self.coef_sampler_type,
            'global_scale_update': self.gscale_update,
            'hmc_curvature_est_stabilized': self.curvature_est_stabilized,
            'noise': self.noise,
            'init_optimizer': self.init_optimizer,
        }

    @staticmethod
    def pick_default_and_create(coef_sampler_type, options, model_name, design):
        """Resolve the sampler (gibbs_util.py:33-84).  Device matrices default to 'cg' (the reference's rule for its
        accelerator mode) and also offer 'cholesky'; 'hmc' is not offered."""
        options = {} if options is None else dict(options)
        if 'coef_sampler_type' in options:
            if coef_sampler_type is not None:
                warn("Duplicate specification of method for sampling "
                     "regression coefficient. Will use the dictionary one.")
            coef_sampler_type = options['coef_sampler_type']
        if coef_sampler_type not in (None, 'cholesky', 'cg', 'hmc'):
            raise ValueError("Unsupported sampler type.")
        if coef_sampler_type not in (None, 'cg', 'cholesky') and getattr(design, 'use_gpu', False):
            raise ValueError("Only the 'cg' and 'cholesky' samplers are supported with device-resident design matrices.")
        if model_name not in ('linear', 'logit'):
            raise ValueError("Only the linear and logit models are supported.")
        n_obs, n_pred = design.shape
        if n_pred > getattr(design, 'n_global', n_obs):
            warn("Sampler has not been optimized for 'small n' problem.")
        options['coef_sampler_type'] = 'cholesky' if coef_sampler_type == 'cholesky' else 'cg'
        return SamplerOptions(**options)


class MarkovChainManager():

    def __init__(self, n_obs, n_pred, n_unshrunk, model_name):
        self.n_obs, self.n_pred, self.n_unshrunk = n_obs, n_pred, n_unshrunk
        self.model_name = model_name
        self._prev_timestamp = None

    @staticmethod
    def sampling_info_keys(sampling_method):
        return ['n_cg_iter'] if sampling_method == 'cg' else []

    get_sampling_info_keys = sampling_info_keys

    def pre_allocate(self, samples, sampling_info, n_post_burnin, thin, params_to_save, sampling_method):
        n_sample = math.floor(n_post_burnin / thin)
        shapes = {
            'coef': (self.n_pred, n_sample),
            'local_scale': (self.n_pred - self.n_unshrunk, n_sample),
            'global_scale': (n_sample,),
            'logp': (n_sample,),
            'obs_prec': (n_sample,) if self.model_name == 'linear' else (self.n_obs, n_sample),
        }
        for key in ('coef', 'local_scale', 'global_scale', 'obs_prec', 'logp'):
            if key in params_to_save:
                # column k is written once per saved iteration: Fortran order makes that write contiguous
                samples[key] = np.zeros(shapes[key], order='F')
        for key in self.sampling_info_keys(sampling_method):
            sampling_info[key] = np.zeros(n_sample)

    @staticmethod
    def _slot(mcmc_iter, n_burnin, thin):
        """Index of the saved sample for this iteration, or None if it is not saved."""
        if mcmc_iter <= n_burnin or (mcmc_iter - n_burnin) % thin != 0:
            return None
        return (mcmc_iter - n_burnin) // thin - 1

    def store_current_state(self, samples, mcmc_iter, n_burnin, thin, coef, lscale,
                            gscale, obs_prec, logp, params_to_save):
        k = self._slot(mcmc_iter, n_burnin, thin)
        if k is None:
            return
        current = {'coef': coef, 'local_scale': lscale, 'global_scale': gscale, 'logp': logp}
        for key, val in current.items():
            if key in params_to_save:
                samples[key][..., k] = val
        if 'obs_prec' in params_to_save:
            samples['obs_prec'][..., k] = obs_prec() if callable(obs_prec) else obs_prec

    def store_sampling_info(self, sampling_info, info, mcmc_iter, n_burnin, thin, sampling_method):
        k = self._slot(mcmc_iter, n_burnin, thin)
        if k is None:
            return
        for key in self.sampling_info_keys(sampling_method):
            sampling_info[key][k] = info[key]

    def merge_outputs(self, prev_samples, prev_mcmc_info, new_samples, new_mcmc_info):
        merged = {key: np.concatenate((prev_samples[key], new_samples[key]), axis=-1) for key in new_samples}
        key = '_reg_coef_sampling_info'
        new_mcmc_info[key] = {
            k: np.concatenate((prev_mcmc_info[key][k], new_mcmc_info[key][k]), axis=-1)
            for k in prev_mcmc_info[key]
        }
        new_mcmc_info['n_iter'] += prev_mcmc_info['n_iter']
        new_mcmc_info['runtime'] += prev_mcmc_info['runtime']
        for k in ('_init_optim_info', 'seed'):
            new_mcmc_info[k] = prev_mcmc_info[k]
        return merged, new_mcmc_info

    def pack_parameters(self, coef, obs_prec, lscale, gscale):
        return {'coef': coef, 'local_scale': lscale, 'global_scale': gscale, 'obs_prec': obs_prec}

    def stamp_time(self, curr_time):
        self._prev_timestamp = curr_time

    def print_status(self, n_status_update, mcmc_iter, n_iter, time_format='minute'):
        if n_status_update == 0:
            return
        if mcmc_iter % int(n_iter / n_status_update) != 0:
            return
        now = time.time()
        elapsed = now - self._prev_timestamp
        if time_format == 'second':
            time_str = "{:.3g} seconds".format(elapsed)
        elif time_format == 'minute':
            time_str = "{:.3g} minutes".format(elapsed / 60)
        else:
            raise ValueError()
        print("{:d} Gibbs iterations complete: {} has elasped since the last update.".format(mcmc_iter, time_str))
        self._prev_timestamp = now
