"""The direct ("cholesky") Gaussian draw (reference: reg_coef_sampler/direct_gaussian_sampler.py:4-44), on the device:
X'WX by libbbgpu's fp64 tensor-core kernel, Jacobi scaling, upper Cholesky factor, the three triangular solves
(bb_cholesky_sample, csrc/bb_chol.cu).  The standard normal vector is drawn on the host from the same numpy stream the
reference uses, so with a common seed the two draws agree to rounding."""
import numpy as np

from .. import _lib


def generate_gaussian_with_weight(design, obs_prec, prior_prec_sqrt, z, rand_gen=None, return_stats=False):
    """
    Generate a multi-variate Gaussian with covariance Sigma
        Sigma^{-1} = X diag(obs_prec) X + diag(prior_prec_sqrt) ** 2
    and mean = Sigma z, where X is the `design` matrix.

    obs_prec : 1-d numpy array, or None to use the precisions resident on the device
    prior_prec_sqrt : 1-d numpy array
    """
    P = design.shape[1]
    if rand_gen is None:
        gaussian_vec = np.random.randn(P)                      # direct_gaussian_sampler.py:26-29
    else:
        gaussian_vec = rand_gen.np_random.randn(P)
    sample = np.empty(P)
    stats = np.zeros(2)
    _lib.check(_lib.load().bb_cholesky_sample(
        design._mat, _lib.dptr(None if obs_prec is None else _lib.as_f64(obs_prec)),
        _lib.dptr(_lib.as_f64(prior_prec_sqrt)), _lib.dptr(_lib.as_f64(z)), _lib.dptr(gaussian_vec),
        _lib.dptr(sample), _lib.dptr(stats)))
    if return_stats:
        return sample, {'fisher_ms': stats[0], 'factorisation_ms': stats[1]}
    return sample
