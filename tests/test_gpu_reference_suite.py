"""GPU: the reference's OWN unit tests of the path's seams, re-run against the device classes.

Mirrors, check for check, tests/test_design_matrix.py (intercept + centring of both design classes, centred Fisher
information full and diagonal, removal of constant columns) and the non-Cox half of tests/test_likelihood_models.py
(gradient and Hessian-vector operators of the linear and logistic models against centred finite differences), with the
same problem sizes, generators (the reference's simulate_data.simulate_design from oracle/_ref) and tolerances
(atol = rtol = 1e-5).  The design matrices are GpuSparseDesignMatrix / GpuDenseDesignMatrix: every product below runs
through libbbgpu.so."""
import numpy as np
import scipy.sparse as sp
import pytest

from conftest import import_reference

pytestmark = pytest.mark.gpu
ATOL = RTOL = 10e-6       # tests/test_design_matrix.py:8-9


def _simulate_design(n_obs, n_pred, binary_frac=0., format_='sparse'):
    """The reference's own generator (simulate_data.simulate_design of the copy under oracle/_ref, which travels to the GPU
    box with the snapshot).  Should that copy be missing, a generator of the same shape takes over (standard-normal columns
    followed by Bernoulli(0.1) columns, from numpy's global stream) -- the checks below do not depend on which one ran."""
    if import_reference() is not None:
        import simulate_data          # oracle/_ref/simulate_data.py (on sys.path after import_reference)
        return simulate_data.simulate_design(n_obs, n_pred, binary_frac=binary_frac, format_=format_)
    n_dense = int(n_pred * (1 - binary_frac))
    X = np.hstack((np.random.randn(n_obs, n_dense), (np.random.rand(n_obs, n_pred - n_dense) < 0.1).astype(float)))
    return sp.csr_matrix(X) if format_ == 'sparse' else X


def _classes():
    from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix, GpuDenseDesignMatrix
    return GpuSparseDesignMatrix, GpuDenseDesignMatrix


def _centred_with_intercept(X):
    X = np.array(X, dtype=float)
    return np.hstack((np.ones((X.shape[0], 1)), X - X.mean(axis=0)))


def test_sparse_design_intercept_and_centering(ctx):          # test_design_matrix.py:12-24
    Sparse, _ = _classes()
    np.random.seed(11)
    X = _simulate_design(100, 10, binary_frac=.5, format_='sparse')
    D = Sparse(X, center_predictor=True, add_intercept=True, ctx=ctx)
    A = _centred_with_intercept(X.toarray())
    w, v = (np.random.randn(size) for size in D.shape)
    assert np.allclose(D.dot(v), A.dot(v), atol=ATOL, rtol=RTOL)
    assert np.allclose(D.Tdot(w), A.T.dot(w), atol=ATOL, rtol=RTOL)


def test_sparse_design_centered_fisher_info(ctx):             # test_design_matrix.py:27-46
    Sparse, _ = _classes()
    # the reference draws the 5 x 3 matrix unseeded; a draw with an all-zero binary column would be cut to 2 columns by
    # both implementations, so take the first seed whose columns all vary
    for seed in range(100):
        np.random.seed(seed)
        X = _simulate_design(5, 3, binary_frac=.5, format_='sparse')
        if np.var(X.toarray(), axis=0).min() > 0:
            break
    D = Sparse(X, center_predictor=True, add_intercept=True, copy_array=True, ctx=ctx)
    A = _centred_with_intercept(X.toarray())
    assert D.shape[1] == A.shape[1]
    weight = np.random.exponential(size=5)
    want = A.T.dot(weight[:, None] * A)
    assert np.allclose(D.compute_fisher_info(weight), want, atol=ATOL, rtol=RTOL)
    assert np.allclose(D.compute_fisher_info(weight, diag_only=True), np.diag(want), atol=ATOL, rtol=RTOL)


def test_dense_design_intercept_and_centering(ctx):           # test_design_matrix.py:49-61
    _, Dense = _classes()
    np.random.seed(13)
    X = _simulate_design(100, 10, binary_frac=.5, format_='dense')
    D = Dense(X.copy(), center_predictor=True, add_intercept=True, ctx=ctx)
    A = _centred_with_intercept(X)
    w, v = (np.random.randn(size) for size in D.shape)
    assert np.allclose(D.dot(v), A.dot(v), atol=ATOL, rtol=RTOL)
    assert np.allclose(D.Tdot(w), A.T.dot(w), atol=ATOL, rtol=RTOL)


def test_intercept_removal(ctx):                              # test_design_matrix.py:71-85
    Sparse, Dense = _classes()
    np.random.seed(14)
    n = 100
    X = _simulate_design(n, 10, binary_frac=.5, format_='sparse')
    with_const = sp.hstack([np.ones((n, 1)), X[:, :5], -.5 * np.ones((n, 1)), X[:, 5:]]).tocsr()
    assert np.allclose(X.toarray(), Sparse.remove_intercept_indicator(with_const).toarray())
    assert np.allclose(X.toarray(), Dense.remove_intercept_indicator(with_const.toarray()))
    # ... and through the constructors (the device computes the column moments): the product sees 10 columns
    for D in (Sparse(with_const, center_predictor=False, add_intercept=False, ctx=ctx),
              Dense(with_const.toarray(), center_predictor=False, add_intercept=False, ctx=ctx)):
        assert D.shape == (n, 10)
        v = np.random.randn(10)
        assert np.allclose(D.dot(v), X.dot(v), atol=ATOL, rtol=RTOL)


# ---- tests/test_likelihood_models.py:12-28 (derivative_tester.py restated: centred differences, dx = 1e-5) ----------
def _gradient_is_close(f, x, dx=10e-6):
    _, grad = f(x)
    est = np.empty(len(x))
    for i in range(len(x)):
        e = np.zeros(len(x)); e[i] = dx
        est[i] = (f(x + e)[0] - f(x - e)[0]) / (2 * dx)
    return np.allclose(grad, est, atol=ATOL, rtol=RTOL)


def _hessian_matvec_is_close(f, x, hess_matvec, n_direction=10, dx=10e-6, seed=0):
    rs = np.random.RandomState(seed)
    for _ in range(n_direction):
        v = rs.randn(len(x))
        v /= np.linalg.norm(v)
        est = (f(x + dx * v)[1] - f(x - dx * v)[1]) / (2 * dx)
        if not np.allclose(hess_matvec(v), est, atol=ATOL, rtol=RTOL):
            return False
    return True


def _simulate_data(model, ctx, n_obs=100, n_pred=50, seed=0):  # tests/helper.py:8-42, design on the device
    from bayesbridge_b200.model import LinearModel, LogisticModel
    Sparse, Dense = _classes()
    np.random.seed(seed)
    X = _simulate_design(n_obs, n_pred, binary_frac=.9)
    beta = np.random.randn(n_pred)
    if model == 'linear':
        y = LinearModel.simulate_outcome(X, beta, noise_sd=1.)
    else:
        n_trial = 1 + np.random.binomial(np.arange(n_obs) + 1, .5)
        y = (LogisticModel.simulate_outcome(n_trial, X, beta), n_trial)
    D = (Sparse(sp.csr_matrix(X), add_intercept=False, ctx=ctx) if sp.issparse(X)
         else Dense(np.ascontiguousarray(X, dtype=float), add_intercept=False, ctx=ctx))
    return y, D, beta


def test_linear_model_gradient_and_hessian(ctx):              # test_likelihood_models.py:12-19
    from bayesbridge_b200.model import LinearModel
    y, D, beta = _simulate_data('linear', ctx)
    model = LinearModel(y, D)
    f = lambda b: model.compute_loglik_and_gradient(b, obs_prec=1.)
    assert _gradient_is_close(f, beta)
    assert _hessian_matvec_is_close(f, beta, model.get_hessian_matvec_operator(beta, 1.))


def test_logistic_model_gradient_and_hessian_matvec(ctx):     # test_likelihood_models.py:22-28 (+ the gradient)
    from bayesbridge_b200.model import LogisticModel
    (n_success, n_trial), D, beta = _simulate_data('logit', ctx)
    model = LogisticModel(n_success, n_trial, D)
    f = model.compute_loglik_and_gradient
    assert _gradient_is_close(f, beta)
    assert _hessian_matvec_is_close(f, beta, model.get_hessian_matvec_operator(beta))
