"""C independent Gibbs chains on ONE dense logit design, advanced in lock-step (BASELINE config 5).

The reference runs one chain per process (bayesbridge.py:109-277); chains that share a design share its most expensive
operand, X.  Here every chain keeps exactly the state and the random streams the single-chain sampler gives it --
`BatchedBayesBridge(...).gibbs(n_iter, seeds=[s_0, ...])` is, chain by chain, `BayesBridge.gibbs(n_iter, seed=s_c)` with
device random numbers -- while the two steps that touch X are batched over the chains on the device:

* beta_c | omega_c, tau_c, lambda_c : `bb_cg_sample_batched`: the CG iterations of all chains in lock-step, the products
  X V and X'(Omega o U) as fp64 tensor-core kernels that read X once for all chains (csrc/bb_batch.cu);
* omega_c | beta_c                  : `bb_pg_from_coef_batched`: one product for the tilts of all chains, Polya-Gamma draws
  on every chain's own Philox stream, the logistic log-likelihoods.

The P-length bookkeeping of a chain (prior scales, running summaries, tau and lambda updates) is the single-chain host
code, run per chain with numpy's global generator switched to that chain's state, so that the scalar Gamma draws are
the ones `gibbs(seed=s_c)` would make."""
import ctypes
import time

import numpy as np

from . import _lib
from .bayesbridge import BayesBridge
from .reg_coef_sampler import SparseRegressionCoefficientSampler


class BatchedBayesBridge:

    def __init__(self, model, prior, n_chains):
        if model.name != 'logit' or model.design.is_sparse:
            raise NotImplementedError("The batched multi-chain sampler covers the logit model on a dense device design.")
        if not 1 <= n_chains <= 16:
            raise ValueError("1 <= n_chains <= 16")
        self.model, self.prior, self.C = model, prior, int(n_chains)
        self.bridges = [BayesBridge(model, prior) for _ in range(self.C)]
        _lib.check(_lib.load().bb_batch_init(model.design._mat, self.C))

    # numpy's global generator is the stream of the scalar Gamma draws (random.py:17-22): one saved state per chain
    def _enter(self, c):
        np.random.set_state(self._np_state[c])

    def _leave(self, c):
        self._np_state[c] = np.random.get_state()

    def gibbs(self, n_iter, n_burnin=0, thin=1, seeds=None, init={'global_scale': 0.1},
              params_to_save=('coef', 'global_scale', 'logp'), n_status_update=0, on_iteration=None):
        """Returns (samples, mcmc_info): every array of `samples` has the chain as its LAST axis
        (samples['coef'] is (P, n_saved, C)); mcmc_info['n_cg_iter'] is (n_iter, C)."""
        lib, design = _lib.load(), self.model.design
        mat, C = design._mat, self.C
        n, P = design.shape
        b0 = self.bridges[0]
        k, bridge_exp = b0.n_unshrunk, self.prior.bridge_exp
        seeds = list(range(C)) if seeds is None else [int(s) for s in seeds]
        if len(seeds) != C:
            raise ValueError("one seed per chain")
        b0._ensure_outcome()
        start = time.time()

        # ---- initialisation: what BayesBridge.gibbs(seed=s_c) does before its loop, per chain; the mode search is a
        # deterministic function of the data and `init`, so it is run once and shared
        self._np_state = [None] * C
        coef = np.empty((C, P)); lscale = np.empty((C, P - k)); gscale = np.empty(C); omega = np.empty((C, n))
        mode = {}
        for c, b in enumerate(self.bridges):
            b.rg.set_seed(seeds[c])
            b.reg_coef_sampler = SparseRegressionCoefficientSampler(
                P, b.prior_sd_for_unshrunk, 'cg', False, self.prior.slab_size)
            if mode:
                b.reg_coef_sampler.search_mode = lambda *a, _m=mode, **kw: (_m['coef'].copy(), dict(_m['info']))
            else:
                found = b.reg_coef_sampler.search_mode

                def remember(*a, _f=found, _m=mode, **kw):
                    cf, info = _f(*a, **kw)
                    _m['coef'], _m['info'] = cf.copy(), dict(info)
                    return cf, info
                b.reg_coef_sampler.search_mode = remember
            cf, _, ls, gs, state, optim_info = b.initialize_chain(dict(init), bridge_exp)
            coef[c], lscale[c], gscale[c], omega[c] = cf, ls, gs, state['obs_prec']
            self._leave(c)
            if c == 0:
                init_optim_info = optim_info
        _lib.check(lib.bb_batch_set_obs_prec(mat, _lib.dptr(np.ascontiguousarray(omega))))
        init_runtime = time.time() - start
        kappa = _lib.as_f64(self.model.n_success - self.model.n_trial / 2)
        z = np.ascontiguousarray(np.tile(design.Tdot(kappa), (C, 1)))        # X' kappa: the same for every chain

        n_saved = (n_iter - n_burnin) // thin
        samples = {}
        if 'coef' in params_to_save:
            samples['coef'] = np.zeros((P, n_saved, C))
        if 'global_scale' in params_to_save:
            samples['global_scale'] = np.zeros((n_saved, C))
        if 'local_scale' in params_to_save:
            samples['local_scale'] = np.zeros((P - k, n_saved, C))
        if 'logp' in params_to_save:
            samples['logp'] = np.zeros((n_saved, C))
        n_cg = np.zeros((n_iter, C), dtype=int)
        pps, x0, s = np.empty((C, P)), np.empty((C, P)), np.empty((C, P))
        seeds_cg, offs_cg = (ctypes.c_uint64 * C)(), (ctypes.c_uint64 * C)()
        seeds_pg, offs_pg = (ctypes.c_uint64 * C)(), (ctypes.c_uint64 * C)()
        n_it, info = (ctypes.c_int * C)(), (ctypes.c_int * C)()
        loglik = np.empty(C)
        atol = 10e-6 * np.sqrt(P)                                           # reg_coef_sampler.py:95
        loop_start = time.time()
        for it in range(1, n_iter + 1):
            if on_iteration is not None:
                on_iteration(it)          # e.g. a benchmark opening its timed region after the warm-up iterations
            # beta | omega, tau, lambda  (reg_coef_sampler.py:60-103), all chains in one batched CG solve
            for c, b in enumerate(self.bridges):
                rcs = b.reg_coef_sampler
                pps[c, :k] = 1 / np.asarray(b.prior_sd_for_unshrunk, dtype=np.float64)
                np.divide(1., rcs.regcoef_summarizer.compute_prior_scale(gscale[c], lscale[c]), out=pps[c, k:])
                x0[c] = rcs.regcoef_summarizer.extrapolate_coef_condmean(gscale[c], lscale[c])
                sd = rcs.regcoef_summarizer.estimate_coef_precond_scale_sd()
                s[c] = rcs.cg_sampler.choose_preconditioner(pps[c], None, design, 'prior', sd)
                seeds_cg[c], offs_cg[c] = b.rg.cg.seed, b.rg.cg._next_offset()
            _lib.check(lib.bb_cg_sample_batched(
                mat, None, _lib.dptr(pps), _lib.dptr(z), _lib.dptr(x0), _lib.dptr(s), float(atol), 500,
                _lib.BB_NOISE_PHILOX, None, None, seeds_cg, offs_cg, _lib.dptr(coef), n_it, info))
            for c, b in enumerate(self.bridges):
                b.reg_coef_sampler.regcoef_summarizer.update(coef[c], gscale[c], lscale[c])
                n_cg[it - 1, c] = n_it[c]
                seeds_pg[c], offs_pg[c] = b.rg.pg.seed, b.rg.pg._next_offset()
            # omega | beta  (bayesbridge.py:397-410) + log-likelihoods, batched; the precisions stay on the device
            _lib.check(lib.bb_pg_from_coef_batched(mat, None, seeds_pg, offs_pg, _lib.dptr(loglik)))
            design.dot_count += int(max(n_it)) + 2
            design.Tdot_count += int(max(n_it)) + 2
            # tau | beta ; lambda | tau, beta ; log posterior: the single-chain code on every chain's own streams
            slot = None
            if it > n_burnin and (it - n_burnin) % thin == 0:
                slot = (it - n_burnin) // thin - 1
            for c, b in enumerate(self.bridges):
                self._enter(c)
                cu = coef[c, k:]
                gscale[c] = b.update_global_scale(gscale[c], cu, bridge_exp, method='sample')
                lscale[c] = b.update_local_scale(gscale[c], cu, bridge_exp)
                self._leave(c)
                if slot is not None:
                    if 'coef' in samples:
                        samples['coef'][:, slot, c] = coef[c]
                    if 'logp' in samples:
                        b._loglik_cache = (coef[c], float(loglik[c]))
                        samples['logp'][slot, c] = b.compute_posterior_logprob(coef[c], gscale[c], None, bridge_exp)
                    g_out, l_out = gscale[c], lscale[c]
                    if self.prior._gscale_paramet == 'coef_magnitude':
                        g_out, l_out = self.prior.adjust_scale(gscale[c], lscale[c].copy(), to='coef_magnitude')
                    if 'global_scale' in samples:
                        samples['global_scale'][slot, c] = g_out
                    if 'local_scale' in samples:
                        samples['local_scale'][:, slot, c] = l_out
        runtime = time.time() - start
        _lib.check(lib.bb_batch_get_obs_prec(mat, _lib.dptr(omega)))
        mcmc_info = {
            'n_chains': C, 'seeds': seeds, 'n_iter': n_iter, 'n_burnin': n_burnin, 'thin': thin,
            'runtime': runtime, 'init_runtime': init_runtime, 'loop_runtime': time.time() - loop_start,
            'n_cg_iter': n_cg, '_init_optim_info': init_optim_info,
            '_markov_chain_state': {'coef': coef.copy(), 'obs_prec': omega, 'local_scale': lscale.copy(),
                                    'global_scale': gscale.copy()},
        }
        return samples, mcmc_info
