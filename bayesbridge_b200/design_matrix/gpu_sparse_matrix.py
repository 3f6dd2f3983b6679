"""Device-resident sparse design matrix (replaces design_matrix/sparse_matrix.py:21-205)."""
import ctypes
import numpy as np
import scipy.sparse as sparse

from .. import _lib
from .abstract_matrix import AbstractDesignMatrix


class GpuSparseDesignMatrix(AbstractDesignMatrix):

    def __init__(self, X, use_mkl=False, center_predictor=False, add_intercept=True,
                 copy_array=False, dot_format='csr', Tdot_format='csr',
                 ctx=None, pattern_only='auto', presharded=False, n_global=None, row_offset=0):
        """
        X : scipy sparse matrix (any format).  With a communicator attached (ctx.nranks > 1) every rank
            passes the full matrix and keeps rows [n*r/G, n*(r+1)/G), or passes its own block with
            presharded=True (+ n_global, row_offset).
        dot_format / Tdot_format : 'csr' or 'csc' -- accepted for API compatibility; the device always
            holds both the CSR image (for dot) and the CSC image (for Tdot).
        pattern_only : 'auto' stores indices only when every stored value equals 1.0.
        """
        super().__init__()
        if not sparse.issparse(X):
            raise TypeError("GpuSparseDesignMatrix expects a scipy sparse matrix.")
        if dot_format not in ('csr', 'csc') or Tdot_format not in ('csr', 'csc'):
            raise NotImplementedError("Unknown sparse format.")
        if copy_array:
            X = X.copy()
        self.ctx = ctx if ctx is not None else _lib.Context.default()
        self.centered = bool(center_predictor)
        self.intercept_added = bool(add_intercept)
        self.use_mkl = False
        sharded = self.ctx.nranks > 1

        if sharded and presharded:
            n_glob = int(n_global)
            X, col_mean = self.remove_intercept_indicator_sharded(X, self.ctx, n_glob)
            X_local = X.tocsr()
        else:
            X = self.remove_intercept_indicator(X)
            n_glob = X.shape[0]
            col_mean = np.squeeze(np.array(X.mean(axis=0))).reshape(-1)
            X_csr = X.tocsr()
            if sharded:
                lo, hi = self.shard_rows(n_glob, self.ctx)
                X_local, row_offset = X_csr[lo:hi], lo
            else:
                X_local = X_csr
        self.column_offset = col_mean if center_predictor else np.zeros(X.shape[1])
        self.X_main = X_local          # host image, kept for toarray() / export checks
        self.n_global = n_glob
        self.row_offset = int(row_offset)

        indptr = np.ascontiguousarray(X_local.indptr)
        indices = np.ascontiguousarray(X_local.indices)
        if indptr.dtype != np.int32 or indices.dtype != np.int32:
            if X_local.nnz >= 2 ** 31 - 1 or max(X_local.shape) >= 2 ** 31 - 1:
                raise ValueError("int64 sparse indices are not supported by the device kernels.")
            indptr, indices = indptr.astype(np.int32), indices.astype(np.int32)
        data = _lib.as_f64(X_local.data)
        if pattern_only == 'auto':
            pattern_only = bool(data.size > 0 and np.all(data == 1.0))
        self.is_binary = bool(pattern_only)
        offset = _lib.as_f64(self.column_offset) if center_predictor else None
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().bb_csr_upload(
            self.ctx.handle, X_local.shape[0], X_local.shape[1], X_local.nnz,
            _lib.iptr(indptr), _lib.iptr(indices), None if self.is_binary else _lib.dptr(data),
            _lib.dptr(offset), int(self.intercept_added), self.row_offset, n_glob, ctypes.byref(handle)))
        self._mat = handle
        self._check_shards_agree()

    @property
    def shape(self):
        n, p = self.X_main.shape
        return n, p + int(self.intercept_added)

    @property
    def is_sparse(self):
        return True

    @property
    def nnz(self):
        return self.X_main.nnz

    def export_csc(self):
        """The CSC image the device built (indptr, indices, data) -- for bit-exactness checks."""
        p, nnz = self.X_main.shape[1], self.X_main.nnz
        indptr, indices, data = np.empty(p + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz)
        _lib.check(_lib.load().bb_mat_export_csc(self._mat, _lib.iptr(indptr), _lib.iptr(indices), _lib.dptr(data)))
        return indptr, indices, data

    def export_csr(self):
        n, nnz = self.X_main.shape[0], self.X_main.nnz
        indptr, indices, data = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz)
        _lib.check(_lib.load().bb_mat_export_csr(self._mat, _lib.iptr(indptr), _lib.iptr(indices), _lib.dptr(data)))
        return indptr, indices, data

    def toarray(self):
        X = self.X_main.toarray() - self.column_offset[np.newaxis, :]
        if self.intercept_added:
            X = np.hstack((np.ones((X.shape[0], 1)), X))
        return X
