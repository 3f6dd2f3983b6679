"""Device-resident dense design matrix (replaces design_matrix/dense_matrix.py:9-70).

The reference materialises [1, X - mean]; the device keeps X raw and applies the intercept and the
centring implicitly, like the sparse class, so one CG pipeline serves both."""
import ctypes
import numpy as np

from .. import _lib
from .abstract_matrix import AbstractDesignMatrix


class GpuDenseDesignMatrix(AbstractDesignMatrix):

    def __init__(self, X, center_predictor=False, add_intercept=True, copy_array=False,
                 ctx=None, presharded=False, n_global=None, row_offset=0):
        super().__init__()
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise TypeError("GpuDenseDesignMatrix expects a 2-d array.")
        self.ctx = ctx if ctx is not None else _lib.Context.default()
        self.centered = bool(center_predictor)
        self.intercept_added = bool(add_intercept)
        sharded = self.ctx.nranks > 1
        if sharded and presharded:
            n_glob = int(n_global)
            X, col_mean = self.remove_intercept_indicator_sharded(X, self.ctx, n_glob)
            X_local = X
        else:
            X = self.remove_intercept_indicator(X)
            n_glob = X.shape[0]
            col_mean = np.mean(X, axis=0)
            if sharded:
                lo, hi = self.shard_rows(n_glob, self.ctx)
                X_local, row_offset = X[lo:hi], lo
            else:
                X_local = X
        self.column_offset = col_mean if center_predictor else np.zeros(X.shape[1])
        self.X_raw = np.ascontiguousarray(X_local)
        self.n_global = n_glob
        self.row_offset = int(row_offset)
        offset = _lib.as_f64(self.column_offset) if center_predictor else None
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().bb_dense_upload(
            self.ctx.handle, self.X_raw.shape[0], self.X_raw.shape[1], _lib.dptr(self.X_raw),
            _lib.dptr(offset), int(self.intercept_added), self.row_offset, n_glob, ctypes.byref(handle)))
        self._mat = handle
        self._check_shards_agree()

    @property
    def shape(self):
        n, p = self.X_raw.shape
        return n, p + int(self.intercept_added)

    @property
    def is_sparse(self):
        return False

    @property
    def nnz(self):
        return self.X_raw.size

    def toarray(self):
        X = self.X_raw - self.column_offset[np.newaxis, :]
        if self.intercept_added:
            X = np.hstack((np.ones((X.shape[0], 1)), X))
        return X

    def extract_matrix(self, order=None):
        return self.toarray()
