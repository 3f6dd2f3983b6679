"""Prior-preconditioned CG sampler, device-resident (replaces reg_coef_sampler/cg_sampler.py:15-151
together with the scipy.sparse.linalg.cg loop it calls)."""
import ctypes
from warnings import warn

import numpy as np

from .. import _lib


class ConjugateGradientSampler():

    def __init__(self, n_coef_wo_shrinkage):
        self.n_coef_wo_shrinkage = n_coef_wo_shrinkage

    def sample(self, design, obs_prec, prior_prec_sqrt, z,
               coef_cg_init=None, precond_by='prior', coef_scaled_sd=None,
               maxiter=None, atol=10e-6, seed=None, noise='host', philox=None, return_stats=False):
        """
        Draw from N(Sigma z, Sigma), Sigma^{-1} = X' diag(obs_prec) X + diag(prior_prec_sqrt)^2.

        Same arguments as the reference, plus
        noise : 'host'   -- eps ~ np.random.randn, drawn exactly as cg_sampler.py:61-62 does and
                            injected into the device solver (the parity mode);
                'device' -- generated inside the kernels from `philox` = (seed, offset).
        obs_prec : array, or None to use the precision vector resident on the device.
        z : array, or None to let the device form X'(obs_prec * y) from the resident outcome.
        """
        n, P = design.shape
        prior_prec_sqrt = _lib.as_f64(prior_prec_sqrt)
        if coef_cg_init is None:
            coef_cg_init = np.zeros(P)
        if seed is not None:
            np.random.seed(seed)
        precond_scale = self.choose_preconditioner(
            prior_prec_sqrt, obs_prec, design, precond_by, coef_scaled_sd)
        if maxiter is None:
            maxiter = 10 * P   # scipy's default
        if noise == 'host':
            # cg_sampler.py:61-62.  On a row-sharded design every rank holds the same seeded numpy stream: draw the
            # n_global normals of the whole problem and keep this shard's rows, then eps2, so that the shards see
            # independent rows of one global eps1 and all ranks consume the stream identically.
            n_global = int(getattr(design, 'n_global', n))
            row_offset = int(getattr(design, 'row_offset', 0))
            if n_global != n:
                eps1 = np.ascontiguousarray(np.random.randn(n_global)[row_offset:row_offset + n])
            else:
                eps1 = np.random.randn(n)
            eps2 = np.random.randn(P)
            mode, sd, off = _lib.BB_NOISE_INJECT, 0, 0
        elif noise == 'device':
            eps1 = eps2 = None
            mode = _lib.BB_NOISE_PHILOX
            sd, off = philox
        else:
            raise ValueError("noise must be 'host' or 'device'")
        coef = np.empty(P)
        n_iter, info = ctypes.c_int(), ctypes.c_int()
        stats = np.zeros(3)
        _lib.check(_lib.load().bb_cg_sample(
            design._mat,
            _lib.dptr(None if obs_prec is None else _lib.as_f64(obs_prec)),
            _lib.dptr(prior_prec_sqrt),
            _lib.dptr(None if z is None else _lib.as_f64(z)),
            _lib.dptr(_lib.as_f64(coef_cg_init)), _lib.dptr(_lib.as_f64(precond_scale)),
            float(atol), int(maxiter), mode, _lib.dptr(eps1), _lib.dptr(eps2),
            int(sd), int(off), _lib.dptr(coef), ctypes.byref(n_iter), ctypes.byref(info), _lib.dptr(stats)))
        design.dot_count += n_iter.value + 1
        design.Tdot_count += n_iter.value + 2
        if info.value != 0:
            warn(
                "The conjugate gradient algorithm did not achieve the requested " +
                "tolerance level. You may increase the maxiter or use the dense " +
                "linear algebra instead."
            )
        cg_info = {'n_iter': n_iter.value, 'valid_input': info.value >= 0, 'converged': info.value == 0}
        if return_stats:
            cg_info.update({'b_norm': stats[0], 'resid_norm': stats[1], 'device_ms': stats[2]})
        return coef, cg_info

    def choose_preconditioner(self, prior_prec_sqrt, obs_prec, design, precond_by, coef_scaled_sd):
        """Diagonal scaling s of cg_sampler.py:123-151."""
        k = self.n_coef_wo_shrinkage
        if precond_by == 'prior':
            scale = np.ones(len(prior_prec_sqrt))
            scale[k:] = prior_prec_sqrt[k:] ** -1
            if k > 0:
                scale[:k] = 2. * np.asarray(coef_scaled_sd)[:k]   # err on the side of large precision
        elif precond_by == 'diag':
            if obs_prec is None:
                raise ValueError("precond_by='diag' needs obs_prec on the host")
            diag = prior_prec_sqrt ** 2 + design.compute_fisher_info(weight=obs_prec, diag_only=True)
            scale = 1 / np.sqrt(diag)
        elif precond_by is None:
            scale = np.ones(design.shape[1])
        else:
            raise NotImplementedError()
        return scale
