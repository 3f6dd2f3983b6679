"""Coefficient update of the Gibbs sampler: the host-side driver of the conditional Gaussian draw
(reference: reg_coef_sampler/reg_coef_sampler.py) and the conjugate-gradient sampler whose linear algebra
runs in libbbgpu.so (reference: reg_coef_sampler/cg_sampler.py)."""
from .reg_coef_sampler import SparseRegressionCoefficientSampler
from .cg_sampler import ConjugateGradientSampler
from .direct_gaussian_sampler import generate_gaussian_with_weight

__all__ = ['SparseRegressionCoefficientSampler', 'ConjugateGradientSampler', 'generate_gaussian_with_weight']
