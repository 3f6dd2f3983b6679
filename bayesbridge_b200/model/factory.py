"""RegressionModel factory (reference: model/factory.py:10-68): picks the device design-matrix class."""
import scipy.sparse

from .linear_model import LinearModel
from .logistic_model import LogisticModel
from ..design_matrix import GpuDenseDesignMatrix, GpuSparseDesignMatrix, AbstractDesignMatrix


def RegressionModel(outcome, X, family='linear', add_intercept=None, center_predictor=True, ctx=None,
                    **design_kwargs):
    """
    outcome : y (linear); n_success or (n_success, n_trial) (logit)
    X : numpy array, scipy sparse matrix, or an already-built device design matrix
    family : 'linear' | 'logit'   ('cox' is outside the CG path and not offered)
    ctx : bayesbridge_b200 Context (default: LOCAL_RANK's GPU)
    """
    if family == 'cox':
        raise NotImplementedError(
            "The Cox model uses the HMC sampler, which is outside the CG path this package implements.")
    if family not in ('linear', 'logit'):
        raise NotImplementedError()
    if add_intercept is None:
        add_intercept = True
    if isinstance(X, AbstractDesignMatrix):
        design = X
    else:
        cls = GpuSparseDesignMatrix if scipy.sparse.issparse(X) else GpuDenseDesignMatrix
        design = cls(X, add_intercept=add_intercept, center_predictor=center_predictor, ctx=ctx, **design_kwargs)

    def local(v):
        # under row sharding the caller passes global outcome vectors; keep this rank's block
        if v is None or len(v) == design.shape[0]:
            return v
        if len(v) == design.n_global:
            return v[design.row_offset:design.row_offset + design.shape[0]]
        raise ValueError("Incompatible sizes of the outcome and design matrix.")

    if family == 'linear':
        return LinearModel(local(outcome), design)
    n_success, n_trial = outcome if isinstance(outcome, tuple) else (outcome, None)
    return LogisticModel(local(n_success), local(n_trial), design)
