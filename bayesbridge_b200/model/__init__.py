"""Likelihood models on device-resident design matrices: Gaussian and logistic (binomial).
The reference's Cox model (model/cox_model.py) is outside this package's scope."""
from .factory import RegressionModel
from .linear_model import LinearModel
from .logistic_model import LogisticModel

__all__ = ['RegressionModel', 'LinearModel', 'LogisticModel']
