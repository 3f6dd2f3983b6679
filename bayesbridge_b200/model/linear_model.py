"""Gaussian likelihood y ~ N(X beta, 1 / obs_prec) on a device-resident design matrix.

Mirrors the interface of the reference's model/linear_model.py:6-45 (same method names, argument meaning and return
values).  Under row sharding `design` holds this rank's rows and `y` the matching outcomes; the scalar sums are
completed over the ranks with `_gsum`, so every rank sees the log-likelihood of the whole data set."""
import math
import numpy as np

from .abstract_model import AbstractModel


class LinearModel(AbstractModel):

    def __init__(self, y, design):
        y = np.asarray(y, dtype=np.float64)
        if len(y) != design.shape[0]:
            raise ValueError("Incompatible sizes of the outcome and design matrix.")
        self.y = y
        self.design = design
        self.name = 'linear'

    def compute_loglik_and_gradient(self, beta, obs_prec, loglik_only=False):
        """Log-likelihood up to the 2 pi constant and its gradient X'(y - X beta) * obs_prec (linear_model.py:13-24)."""
        if getattr(self.design, '_mat', None) is not None:
            ll, grad = self._device_loglik_and_gradient(beta, float(obs_prec), loglik_only)
            return self.n_obs_global * math.log(obs_prec) / 2 + ll, grad
        resid = self.y - self.design.dot(beta)
        loglik = self.n_obs_global * math.log(obs_prec) / 2 - obs_prec * self._gsum(np.sum(resid ** 2)) / 2
        grad = None if loglik_only else obs_prec * self.design.Tdot(resid)
        return loglik, grad

    def _outcome_arrays(self):
        return None, self.y

    def get_hessian_matvec_operator(self, beta, obs_prec):
        """v -> -obs_prec X'X v, two device products per application (linear_model.py:29-31)."""
        return lambda v: - obs_prec * self.design.Tdot(self.design.dot(v))

    def calc_intercept_mle(self):
        """Mean outcome over all ranks (linear_model.py:33-34)."""
        return self._gsum(self.y.sum()) / self.n_obs_global

    @staticmethod
    def simulate_outcome(X, beta, noise_sd, seed=None):
        """X beta + Gaussian noise from numpy's global stream; X only needs `dot` (linear_model.py:36-45)."""
        if seed is not None:
            np.random.seed(seed)
        return X.dot(beta) + noise_sd * np.random.randn(X.shape[0])
