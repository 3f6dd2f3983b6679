"""Gaussian likelihood (reference: model/linear_model.py)."""
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
        resid = self.y - self.design.dot(beta)
        loglik = self.n_obs_global * math.log(obs_prec) / 2 - obs_prec * self._gsum(np.sum(resid ** 2)) / 2
        grad = None if loglik_only else obs_prec * self.design.Tdot(resid)
        return loglik, grad

    def get_hessian_matvec_operator(self, beta, obs_prec):
        return lambda v: - obs_prec * self.design.Tdot(self.design.dot(v))

    def calc_intercept_mle(self):
        return self._gsum(self.y.sum()) / self.n_obs_global

    @staticmethod
    def simulate_outcome(X, beta, noise_sd, seed=None):
        if seed is not None:
            np.random.seed(seed)
        return X.dot(beta) + noise_sd * np.random.randn(X.shape[0])
