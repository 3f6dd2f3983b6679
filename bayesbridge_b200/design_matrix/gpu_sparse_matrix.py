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

        # Row block of this rank.  The column moments that decide which columns are constant (abstract_matrix.py:93-107)
        # and give the centring offsets are taken on the DEVICE from the CSC image the upload builds anyway
        # (bb_column_moments), instead of two host passes over the matrix (X.mean, X.power(2).mean: seconds at nnz = 1e8);
        # only when a constant column is found (rare: a hand-added intercept) is the matrix cut on the host and sent again.
        import time
        t0 = time.perf_counter()
        X_csr = X.tocsr()
        if sharded and not presharded:
            n_glob = X_csr.shape[0]
            lo, hi = self.shard_rows(n_glob, self.ctx)
            X_local, row_offset = X_csr[lo:hi], lo
        else:
            n_glob = int(n_global) if (sharded and presharded) else X_csr.shape[0]
            X_local = X_csr
        self.n_global = n_glob
        self.row_offset = int(row_offset)
        # columns without a single stored entry anywhere (rare features of a sparse binary design) are constant by
        # construction: one counting pass finds them, so that the matrix is cut BEFORE it is uploaded
        if X_local.shape[1] > 0:
            present = np.bincount(X_local.indices, minlength=X_local.shape[1]).astype(np.float64)
            if sharded:
                present = self.ctx.allreduce_host(present)
            if np.any(present == 0):
                import warnings
                warnings.warn(
                    "Intercept column (or numerically indistinguishable from such) detected. "
                    "Do not add intercept manually. Removing....")
                X_local = X_local[:, present > 0].tocsr()
        t_prep = time.perf_counter() - t0
        t0 = time.perf_counter()
        self._mat = None
        self._upload(X_local, pattern_only, n_glob)
        mean, constant = self._device_column_moments(X_local.shape[1], n_glob)
        if np.any(constant):
            import warnings
            warnings.warn(
                "Intercept column (or numerically indistinguishable from such) detected. "
                "Do not add intercept manually. Removing....")
            keep = np.logical_not(constant)
            X_local, mean = X_local[:, keep].tocsr(), mean[keep]
            _lib.check(_lib.load().bb_mat_free(self._mat))
            self._mat = None
            self._upload(X_local, pattern_only, n_glob)
        self.column_offset = mean if center_predictor else np.zeros(X_local.shape[1])
        if center_predictor:
            _lib.check(_lib.load().bb_set_column_offset(self._mat, _lib.dptr(_lib.as_f64(self.column_offset))))
        self.X_main = X_local          # host image, kept for toarray() / export checks
        self.build_seconds = {'host_prepare': t_prep, 'upload_and_moments': time.perf_counter() - t0}
        self._check_shards_agree()

    def _upload(self, X_local, pattern_only, n_glob):
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
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().bb_csr_upload(
            self.ctx.handle, X_local.shape[0], X_local.shape[1], X_local.nnz,
            _lib.iptr(indptr), _lib.iptr(indices), None if self.is_binary else _lib.dptr(data),
            None, int(self.intercept_added), self.row_offset, n_glob, ctypes.byref(handle)))
        self._mat = handle

    def _device_column_moments(self, p, n_glob):
        """(column means, constant-column mask) from the device's column sums / sums of squares, summed over the shards."""
        s1, s2 = np.zeros(p), np.zeros(p)
        _lib.check(_lib.load().bb_column_moments(self._mat, _lib.dptr(s1), _lib.dptr(s2)))
        if self.ctx.nranks > 1:
            tot = self.ctx.allreduce_host(np.concatenate((s1, s2)))
            s1, s2 = tot[:p], tot[p:]
        mean, sq_mean = s1 / n_glob, s2 / n_glob
        constant = (sq_mean - mean ** 2) < n_glob * 2.0 ** -52
        return mean, constant


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
