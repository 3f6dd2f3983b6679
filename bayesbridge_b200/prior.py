"""Bridge prior and its hyper-parameters (reference: prior.py). Scalar host maths."""
import math
from warnings import warn

import numpy as np
import scipy.optimize
from scipy.special import polygamma


class RegressionCoefPrior():

    def __init__(self, bridge_exponent=.5, n_fixed_effect=0, sd_for_intercept=float('inf'),
                 sd_for_fixed_effect=float('inf'), regularizing_slab_size=float('inf'),
                 global_scale_prior_hyper_param=None, _global_scale_parametrization='coef_magnitude'):
        """
        bridge_exponent : exponent (< 2) of the bridge prior exp(-|beta/tau|^alpha)
        n_fixed_effect : leading predictors given Gaussian priors instead of shrinkage
        sd_for_intercept, sd_for_fixed_effect : prior sds (inf = flat)
        regularizing_slab_size : sd of the Gaussian slab that bounds the bridge tails
        global_scale_prior_hyper_param : None, or {'log10_mean', 'log10_sd'} of log10(global scale)
        """
        if not (np.isscalar(sd_for_fixed_effect) or n_fixed_effect == len(sd_for_fixed_effect)):
            raise ValueError("Prior sd for fixed effects must be a scalar or have length n_fixed_effect.")
        if bridge_exponent > 2:
            raise ValueError("Exponent larger than 2 is unsupported.")
        if np.isscalar(sd_for_fixed_effect):
            sd_for_fixed_effect = sd_for_fixed_effect * np.ones(n_fixed_effect)
        self.sd_for_intercept = sd_for_intercept
        self.sd_for_fixed = sd_for_fixed_effect
        self.slab_size = regularizing_slab_size
        self.n_fixed = n_fixed_effect
        self.bridge_exp = bridge_exponent
        self._gscale_paramet = _global_scale_parametrization
        if global_scale_prior_hyper_param is None:
            # reference (scale-invariant) prior
            self.param = {'gscale_neg_power': {'shape': 0., 'rate': 0.}, 'gscale': None}
        else:
            if not ({'log10_mean', 'log10_sd'} <= global_scale_prior_hyper_param.keys()):
                raise ValueError("Dictionary should contain keys 'log10_mean' and 'log10_sd.'")
            log10_mean = global_scale_prior_hyper_param['log10_mean']
            log10_sd = global_scale_prior_hyper_param['log10_sd']
            shape, rate = self.solve_for_gscale_prior_hyperparam(
                log10_mean, log10_sd, bridge_exponent, self._gscale_paramet)
            self.param = {
                'gscale_neg_power': {'shape': shape, 'rate': rate},   # in the 'raw' parametrisation
                'gscale': {'log10_mean': log10_mean, 'log10_sd': log10_sd},
            }

    def get_info(self):
        sd_fixed = self.sd_for_fixed
        if len(sd_fixed) > 0 and np.all(sd_fixed == sd_fixed[0]):
            sd_fixed = sd_fixed[0]
        return {
            'bridge_exponent': self.bridge_exp,
            'n_fixed_effect': self.n_fixed,
            'sd_for_intercept': self.sd_for_intercept,
            'sd_for_fixed_effect': sd_fixed,
            'regularizing_slab_size': self.slab_size,
            'global_scale_prior_hyper_param': self.param['gscale'],
            '_global_scale_parametrization': self._gscale_paramet,
        }

    def clone(self, **kwargs):
        """Copy with the given constructor arguments replaced."""
        info = self.get_info()
        if '_global_scale_parametrization' in kwargs:
            raise ValueError("Change of parametrization is not supported.")
        for key, val in kwargs.items():
            if key in info:
                info[key] = val
            else:
                warn("'{:s} is not a valid keyward argument.".format(key))
        return RegressionCoefPrior(**info)

    def adjust_scale(self, gscale, lscale, to):
        """Move (tau, lambda) between the raw and the coefficient-magnitude parametrisations."""
        unit = self.compute_power_exp_ave_magnitude(self.bridge_exp, 1.)
        if to == 'raw':
            gscale /= unit
            lscale *= unit
        elif to == 'coef_magnitude':
            gscale *= unit
            lscale /= unit
        else:
            raise ValueError()
        return gscale, lscale

    @staticmethod
    def compute_power_exp_ave_magnitude(exponent, scale=1.):
        """E|x| under the density proportional to exp(-|x/scale|^exponent)."""
        return scale * math.gamma(2 / exponent) / math.gamma(1 / exponent)

    @staticmethod
    def change_log_base(val, from_=math.e, to=10.):
        return val * math.log(from_) / math.log(to)

    def solve_for_gscale_prior_hyperparam(self, log10_mean, log10_sd, bridge_exp, gscale_paramet):
        log_mean = self.change_log_base(log10_mean, from_=10., to=math.e)
        log_sd = self.change_log_base(log10_sd, from_=10., to=math.e)
        if gscale_paramet == 'coef_magnitude':
            log_mean -= math.log(self.compute_power_exp_ave_magnitude(bridge_exp, 1.))
        return self.solve_for_gamma_param(log_mean, log_sd, bridge_exp)

    def solve_for_gamma_param(self, log_mean, log_sd, bridge_exp):
        """Gamma(shape, rate) prior on phi = tau^(-alpha) whose implied log(tau) has the given mean / sd:
        sd(log phi) = sqrt(trigamma(shape)), E log phi = digamma(shape) - log(rate)."""
        if log_sd < 0:
            raise ValueError("Variance has to be positive.")
        if log_sd > 10 ** 8:
            raise ValueError("Specified prior variance is too large.")

        def f(log_shape):
            return math.sqrt(self._polygamma(1, math.exp(log_shape))) / bridge_exp - log_sd

        lower, upper = self._find_root_bounds(f, -10.)
        log_shape = scipy.optimize.brentq(f, lower, upper)
        shape = math.exp(log_shape)
        rate = math.exp(self._polygamma(0, shape) + bridge_exp * log_mean)
        return shape, rate

    @staticmethod
    def _polygamma(n, x):
        return polygamma([n], x)[0]

    @staticmethod
    def _find_root_bounds(f, init_lower_lim, increment=5., max_lim=None):
        if max_lim is None:
            max_lim = init_lower_lim + 10 ** 4
        if f(init_lower_lim) < 0:
            raise ValueError("Objective function must have positive value at the lower limit.")
        lo = init_lower_lim
        while f(lo + increment) > 0 and lo < max_lim:
            lo += increment
        if lo >= max_lim:
            raise Exception("could not bracket the root")
        return lo, lo + increment
