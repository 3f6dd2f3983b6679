class AbstractModel():
    """What BayesBridge needs from a likelihood (reference: model/abstract_model.py)."""

    @property
    def n_obs(self):
        return self.design.shape[0]

    @property
    def n_pred(self):
        return self.design.shape[1]

    @property
    def intercept_added(self):
        return self.design.intercept_added

    # ---- device-resident evaluation of the log-likelihood and its gradient (libbbgpu: bb_loglik_and_gradient) ----
    def _outcome_arrays(self):
        raise NotImplementedError()

    def _ensure_outcome_resident(self):
        """The outcome lives next to X on the device handle of the design; a design may serve several models."""
        from .. import _lib
        design = self.design
        if getattr(design, '_outcome_model', None) != id(self):
            n_trial, y = self._outcome_arrays()
            _lib.check(_lib.load().bb_set_outcome(design._mat, _lib.dptr(None if n_trial is None else _lib.as_f64(n_trial)),
                                                  _lib.dptr(_lib.as_f64(y))))
            design._outcome_model = id(self)
            design._outcome_owner = None        # a bridge bound to this design re-checks its own state

    def _device_loglik_and_gradient(self, beta, obs_prec, loglik_only):
        """One device pass: eta = X beta, the likelihood terms and X'residual; only P-vectors cross PCIe.  The last
        evaluation is kept, because L-BFGS asks for the value and the gradient of the same point in separate calls."""
        import ctypes
        import numpy as np
        from .. import _lib
        beta = _lib.as_f64(beta)
        key = (beta.tobytes(), float(obs_prec))
        cached = getattr(self, '_ll_cache', None)
        if cached is not None and cached[0] == key and (loglik_only or cached[2] is not None):
            return cached[1], (None if loglik_only else cached[2].copy())
        self._ensure_outcome_resident()
        ll = _lib.c_dbl()
        grad = None if loglik_only else np.empty(self.design.shape[1])
        _lib.check(_lib.load().bb_loglik_and_gradient(self.design._mat, _lib.dptr(beta), float(obs_prec), int(bool(loglik_only)),
                                                      ctypes.byref(ll), _lib.dptr(grad)))
        self.design.dot_count += 1
        if not loglik_only:
            self.design.Tdot_count += 1
        self._ll_cache = (key, ll.value, None if grad is None else grad.copy())
        return ll.value, grad

    def _gsum(self, x):
        """Sum of a per-observation quantity over ALL row shards (identity without a communicator)."""
        ctx = getattr(self.design, 'ctx', None)
        total = float(x)
        if ctx is not None and ctx.nranks > 1:
            import numpy as np
            total = float(ctx.allreduce_host(np.array([total]))[0])
        return total

    @property
    def n_obs_global(self):
        return getattr(self.design, 'n_global', self.design.shape[0])
