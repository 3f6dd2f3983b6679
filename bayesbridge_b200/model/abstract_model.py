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
