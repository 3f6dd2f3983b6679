"""Duck type every design matrix implements (reference: design_matrix/abstract_matrix.py:14-107)."""
import abc
import warnings
import ctypes
import numpy as np
import scipy.sparse as sparse

from .. import _lib


class AbstractDesignMatrix(abc.ABC):

    use_cupy = False   # read by SamplerOptions in the reference (gibbs_util.py:49-57)
    use_gpu = True     # the B200 classes: products run in libbbgpu, only 'cg' sampling is offered

    def __init__(self):
        self.dot_count = 0
        self.Tdot_count = 0
        self.memoized = False
        self.X_dot_v = None
        self.v_prev = None
        self._mat = None
        self.ctx = None

    # ---- interface -----------------------------------------------------------------------
    @property
    @abc.abstractmethod
    def shape(self):
        """(local rows, p + intercept)"""

    @property
    @abc.abstractmethod
    def is_sparse(self):
        pass

    @abc.abstractmethod
    def toarray(self):
        pass

    # ---- products through the C-ABI (seam 1) ----------------------------------------------
    def dot(self, v):
        """X v with the implicit intercept column and centring (sparse_matrix.py:68-101)."""
        v = _lib.as_f64(v)
        if v.shape != (self.shape[1],):
            raise ValueError("dot: expected a vector of length {}".format(self.shape[1]))
        if self.memoized:
            if np.all(self.v_prev == v):
                return self.X_dot_v
            self.v_prev = v.copy()
        out = np.empty(self.shape[0])
        _lib.check(_lib.load().bb_dot(self._mat, _lib.dptr(v), _lib.dptr(out)))
        if self.memoized:
            self.X_dot_v = out
        self.dot_count += 1
        return out

    def Tdot(self, w):
        """X' w (sparse_matrix.py:103-129); summed over row shards when sharded."""
        w = _lib.as_f64(w)
        if w.shape != (self.shape[0],):
            raise ValueError("Tdot: expected a vector of length {}".format(self.shape[0]))
        out = np.empty(self.shape[1])
        _lib.check(_lib.load().bb_tdot(self._mat, _lib.dptr(w), _lib.dptr(out)))
        self.Tdot_count += 1
        return out

    def compute_fisher_info(self, weight, diag_only=False):
        """X' W X: its diagonal (sparse_matrix.py:164-177) or the full P x P matrix (sparse_matrix.py:131-162,
        dense_matrix.py:54-58) -- the latter by the fp64 tensor-core tile kernel of libbbgpu (bb_chol.cu); a sparse
        design is densified on the device for it, so it is meant for the p <~ 1e4 the Cholesky sampler is for."""
        weight = _lib.as_f64(weight)
        if weight.shape != (self.shape[0],):
            raise ValueError("weight must have one entry per observation")
        if not diag_only:
            P = self.shape[1]
            out = np.empty((P, P))
            _lib.check(_lib.load().bb_fisher_full(self._mat, _lib.dptr(weight), _lib.dptr(out), None))
            return out
        out = np.empty(self.shape[1])
        _lib.check(_lib.load().bb_fisher_diag(self._mat, _lib.dptr(weight), _lib.dptr(out)))
        return out

    def compute_transposed_fisher_info(self, weight, include_intrcpt=False):
        raise NotImplementedError("Only needed by the Cox model, which is out of scope here.")

    # ---- bookkeeping identical to the reference ---------------------------------------------
    def memoize_dot(self, flag=True):
        self.memoized = flag
        if self.v_prev is None:
            self.v_prev = np.full(self.shape[1], float('nan'))
        if not flag:
            self.X_dot_v = None
            self.v_prev = None

    @property
    def n_matvec(self):
        return self.dot_count + self.Tdot_count

    def get_dot_count(self):
        return self.dot_count, self.Tdot_count

    def reset_matvec_count(self, count=0):
        if not hasattr(count, "__len__"):
            count = 2 * [count]
        self.dot_count, self.Tdot_count = count[0], count[1]

    # ---- helpers ---------------------------------------------------------------------------
    @staticmethod
    def is_cupy_matrix(X):
        return False

    is_cupy_dense = is_cupy_sparse = is_cupy_matrix

    @staticmethod
    def remove_intercept_indicator(X):
        """Drop columns whose variance is numerically zero (abstract_matrix.py:93-107)."""
        if sparse.issparse(X):
            mean = np.asarray(X.mean(axis=0)).ravel()
            sq_mean = np.asarray(X.power(2).mean(axis=0)).ravel()
            col_var = sq_mean - mean ** 2
        else:
            col_var = np.var(X, axis=0)
        constant = col_var < X.shape[0] * 2.0 ** -52
        if np.any(constant):
            warnings.warn(
                "Intercept column (or numerically indistinguishable from such) detected. "
                "Do not add intercept manually. Removing....")
            X = X[:, np.logical_not(constant)]
        return X

    @staticmethod
    def remove_intercept_indicator_sharded(X_local, ctx, n_global):
        """Same rule as remove_intercept_indicator, for a matrix given as per-rank row blocks: the column
        moments are summed over all ranks first, so that every rank drops the SAME columns (a column that is
        constant inside one block only must stay, or the ranks would disagree on p and the allreduce would hang)."""
        if sparse.issparse(X_local):
            s1 = np.asarray(X_local.sum(axis=0)).ravel()
            s2 = np.asarray(X_local.power(2).sum(axis=0)).ravel()
        else:
            s1, s2 = X_local.sum(axis=0), (X_local ** 2).sum(axis=0)
        tot = ctx.allreduce_host(np.concatenate((s1, s2)))
        p = X_local.shape[1]
        mean, sq_mean = tot[:p] / n_global, tot[p:] / n_global
        constant = (sq_mean - mean ** 2) < n_global * 2.0 ** -52
        if np.any(constant):
            warnings.warn(
                "Intercept column (or numerically indistinguishable from such) detected. "
                "Do not add intercept manually. Removing....")
            X_local = X_local[:, np.logical_not(constant)]
            mean = mean[np.logical_not(constant)]
        return X_local, mean

    def _check_shards_agree(self):
        """Every rank must hold the same number of columns, or the allreduces inside libbbgpu would hang."""
        if self.ctx.nranks > 1:
            P = float(self.shape[1])
            tot = self.ctx.allreduce_host(np.array([P, P * P]))
            if tot[0] != self.ctx.nranks * P or tot[1] != self.ctx.nranks * P * P:
                raise ValueError("Row shards disagree on the number of predictors.")
            # (p+1)-vectors are the per-CG-iteration exchange: give them the peer-memory all-reduce
            self.ctx.init_p2p(int(P) + 64)

    @staticmethod
    def shard_rows(n, ctx):
        """Contiguous row block of this rank: [lo, hi)."""
        G, r = ctx.nranks, ctx.rank
        return (n * r) // G, (n * (r + 1)) // G

    def info(self):
        lib = _lib.load()
        n, P, nnz = _lib.c_i64(), _lib.c_i64(), _lib.c_i64()
        sp_, bi = _lib.c_int(), _lib.c_int()
        _lib.check(lib.bb_mat_info(self._mat, ctypes.byref(n), ctypes.byref(P), ctypes.byref(nnz),
                                   ctypes.byref(sp_), ctypes.byref(bi)))
        return {'n_local': n.value, 'P': P.value, 'nnz': nnz.value, 'is_sparse': bool(sp_.value),
                'is_binary': bool(bi.value)}

    def time_kernel(self, what, reps=20, flush_l2=True):
        """Mean device milliseconds of one kernel class ('dot' | 'tdot' | 'op') on resident data."""
        ms = _lib.c_dbl()
        _lib.check(_lib.load().bb_time_kernel(self._mat, what.encode(), int(reps), int(bool(flush_l2)), ctypes.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self._mat is not None:
                _lib.load().bb_mat_free(self._mat)
                self._mat = None
        except Exception:
            pass
