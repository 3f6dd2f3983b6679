"""Binomial-logit likelihood (reference: model/logistic_model.py)."""
from warnings import warn
import numpy as np

from .abstract_model import AbstractModel


class LogisticModel(AbstractModel):

    def __init__(self, n_success, n_trial, design):
        n_success = np.asarray(n_success)
        if n_trial is None:
            if np.max(n_success) > 1:
                raise ValueError("If not binary, the number of trials must be specified.")
            if len(n_success) != design.shape[0]:
                raise ValueError("Incompatible sizes of the outcome and design matrix.")
            warn("The numbers of trials were not specified. The binary outcome is assumed.")
            n_trial = np.ones(len(n_success))
        else:
            n_trial = np.asarray(n_trial)
            if not (len(n_trial) == len(n_success) == design.shape[0]):
                raise ValueError("Incompatible sizes of the outcome vectors and design matrix.")
            if np.any(n_trial <= 0):
                raise ValueError("Number of trials must be strictly positive.")
            if np.any(n_success > n_trial):
                raise ValueError("Number of successes cannot be larger than that of trials.")
        self.n_trial = n_trial.astype('float64')
        self.n_success = n_success.astype('float64')
        self.design = design
        self.name = 'logit'

    def _outcome_arrays(self):
        return self.n_trial, self.n_success

    def compute_loglik_and_gradient(self, beta, loglik_only=False):
        """logistic_model.py:49-55.  With a device-resident design the whole evaluation runs there."""
        if getattr(self.design, '_mat', None) is not None:
            return self._device_loglik_and_gradient(beta, 1.0, loglik_only)
        eta = self.design.dot(beta)
        loglik = self._gsum(np.sum(self.n_success * eta - self.n_trial * np.logaddexp(0, eta)))
        if loglik_only:
            return loglik, None
        prob = self.convert_to_probability_scale(eta)
        return loglik, self.design.Tdot(self.n_success - self.n_trial * prob)

    def get_hessian_matvec_operator(self, beta):
        prob = self.compute_predicted_prob(self.design, beta)
        weight = self.n_trial * prob * (1 - prob)
        return lambda v: - self.design.Tdot(weight * self.design.dot(v))

    def calc_intercept_mle(self):
        p_hat = self._gsum(self.n_success.sum()) / self._gsum(self.n_trial.sum())
        return np.log(p_hat / (1 - p_hat))

    @staticmethod
    def compute_polya_gamma_mean(shape, tilt):
        """E[PG(b, c)] = b/(2c) tanh(c/2), with the c -> 0 limit b/4 (logistic_model.py:80-87)."""
        pg_mean = shape.copy() / 2
        nz = np.abs(tilt) > 1e-5
        pg_mean[nz] *= 1 / tilt[nz] * (np.exp(tilt[nz]) - 1) / (np.exp(tilt[nz]) + 1)
        return pg_mean

    @staticmethod
    def compute_predicted_prob(X, beta, truncate=False):
        return LogisticModel.convert_to_probability_scale(X.dot(beta), truncate)

    @staticmethod
    def convert_to_probability_scale(logit_prob, truncate=False):
        if truncate:
            logit_prob = np.clip(logit_prob, -709., 36.7)
        return 1 / (1 + np.exp(-logit_prob))

    @staticmethod
    def simulate_outcome(n_trial, X, beta, seed=None):
        prob = LogisticModel.compute_predicted_prob(X, beta)
        if seed is not None:
            np.random.seed(seed)
        return np.random.binomial(n_trial, prob)
