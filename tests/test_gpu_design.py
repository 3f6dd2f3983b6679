"""GPU: design-matrix products through the C-ABI vs the oracle and vs outputs of the reference
(golden/design_ref.npz); bit-exact CSR/CSC construction; edge cases."""
import numpy as np
import scipy.sparse as sp
import pytest

from conftest import golden
from oracle import cg_oracle as co

pytestmark = pytest.mark.gpu


def _designs():
    from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix, GpuDenseDesignMatrix
    return GpuSparseDesignMatrix, GpuDenseDesignMatrix


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def random_sparse(n, p, density, seed, binary=False, hot_column=True):
    rs = np.random.RandomState(seed)
    X = sp.random(n, p, density=density, format='csr', random_state=rs, dtype=np.float64)
    if hot_column and n > 10 and p > 3:      # a column with n/2 entries: segments that straddle many tiles
        rows = rs.choice(n, n // 2, replace=False)
        X = (X + sp.csr_matrix((np.ones(n // 2), (rows, np.full(n // 2, 2))), shape=(n, p))).tocsr()
    if binary:
        X.data[:] = 1.0
    return X


def test_products_match_reference_outputs(ctx):
    Sparse, Dense = _designs()
    g = golden('design_ref.npz')
    Xd = g['X_dense_image']
    for c in (0, 1):
        for i in (0, 1):
            key = '%d%d' % (c, i)
            for D, pre in ((Sparse(sp.csr_matrix(Xd), center_predictor=bool(c), add_intercept=bool(i), ctx=ctx), ''),
                           (Dense(Xd.copy(), center_predictor=bool(c), add_intercept=bool(i), ctx=ctx), 'd')):
                assert relerr(D.dot(g['v_' + key]), g[pre + 'dot_' + key]) < 1e-13
                assert relerr(D.Tdot(g['w']), g[pre + 'tdot_' + key]) < 1e-13
                assert relerr(D.compute_fisher_info(g['weight'], diag_only=True), g[pre + 'fisher_' + key]) < 1e-12


@pytest.mark.parametrize('n,p,density', [(300, 40, 0.2), (20000, 700, 0.02), (60000, 3000, 0.004)])
@pytest.mark.parametrize('binary', [False, True])
@pytest.mark.parametrize('slab,stage,variant,permute', [(0, 1, 1, 2), (64, 1, 1, 2), (1024, 1, 1, 2), (0, 1, 1, 1), (1024, 1, 1, 0), (0, 1, 2, 2),
                                                        (0, 0, 0, 1), (0, 1, 0, 1), (64, 1, 0, 1), (1024, 1, 0, 1), (1024, 0, 0, 1)])
def test_sparse_products_vs_oracle(ctx, n, p, density, binary, slab, stage, variant, permute):
    """variant 1 = sliced lane-per-fragment kernel (bb_sell.cu, the default), 0 = tile + segmented-scan kernel,
    2 = sub-warp-per-segment kernel on the canonical CSR / CSC image (small, L2-resident problems);
    permute = bank-aware entry order (2: most-loaded-bank-first matching, 1: greedy, 0: canonical order)."""
    Sparse, _ = _designs()
    default_permute = ctx.get_option('bank_permute')
    ctx.set_option('slab_width', slab)
    ctx.set_option('spmv_stage', stage)
    ctx.set_option('spmv_variant', variant)
    ctx.set_option('bank_permute', permute)
    try:
        X = random_sparse(n, p, density, seed=n + p, binary=binary)
        rng = np.random.default_rng(5)
        for center in (False, True):
            for icpt in (False, True):
                D = Sparse(X, center_predictor=center, add_intercept=icpt, ctx=ctx)
                assert D.is_binary == binary
                O = co.DesignOracle(X, center, icpt)
                v, w, wt = rng.standard_normal(D.shape[1]), rng.standard_normal(n), rng.random(n)
                assert relerr(D.dot(v), O.dot(v)) < 1e-12
                assert relerr(D.Tdot(w), O.Tdot(w)) < 1e-12
                assert relerr(D.compute_fisher_info(wt, diag_only=True), O.fisher_diag(wt)) < 1e-12
    finally:
        ctx.set_option('slab_width', 0)
        ctx.set_option('spmv_stage', 1)
        ctx.set_option('spmv_variant', 1)
        ctx.set_option('bank_permute', default_permute)


def test_products_are_bit_reproducible(ctx):
    Sparse, _ = _designs()
    X = random_sparse(30000, 900, 0.01, seed=3)
    D = Sparse(X, center_predictor=True, add_intercept=True, ctx=ctx)
    rng = np.random.default_rng(0)
    v, w = rng.standard_normal(901), rng.standard_normal(30000)
    assert np.array_equal(D.dot(v), D.dot(v))
    assert np.array_equal(D.Tdot(w), D.Tdot(w))


def test_bank_aware_order_changes_rounding_only(ctx):
    """The nnz order inside the slab copies (option bank_permute) is a storage detail: the canonical
    CSR / CSC images stay bit-exact and the products move by rounding only."""
    Sparse, _ = _designs()
    X = random_sparse(60000, 3000, 0.004, seed=11)
    w = np.random.default_rng(2).standard_normal(60000)
    out = {}
    ctx.set_option('slab_width', 1024)
    default_permute = ctx.get_option('bank_permute')
    try:
        for perm in (2, 1, 0):
            ctx.set_option('bank_permute', perm)
            D = Sparse(X, center_predictor=True, add_intercept=True, ctx=ctx, pattern_only=False)
            v = np.random.default_rng(3).standard_normal(D.shape[1])
            out[perm] = (D.dot(v), D.Tdot(w), D.export_csr() + D.export_csc())
    finally:
        ctx.set_option('bank_permute', default_permute)
        ctx.set_option('slab_width', 0)
    for perm in (1, 2):
        assert relerr(out[perm][0], out[0][0]) < 1e-14 and relerr(out[perm][1], out[0][1]) < 1e-14
        for a, b in zip(out[perm][2], out[0][2]):
            assert np.array_equal(a, b)
    C = X.tocsc()
    assert np.array_equal(out[1][2][4], C.indices) and np.array_equal(out[1][2][5], C.data)


@pytest.mark.parametrize('slab', [0, 256, 2048])
def test_work_partition_does_not_change_a_bit(ctx, slab):
    """The sliced kernel's work partition (option sell_partition: 1 = equal-cost CTA ranges that may straddle slabs,
    2 = slab-aligned ranges, 0 = whichever has the shorter estimated critical path), its cost model (sell_slice_cost) and
    the way a section is split into warp strips (sell_lpt: longest slice first + strip-major layout, or contiguous cuts)
    only decide WHO sums a fragment and WHERE its indices are stored: every product must come out bit-identical, and equal to the oracle's."""
    Sparse, _ = _designs()
    X = random_sparse(50000, 3000, 0.004, 5, binary=True)
    O = co.DesignOracle(X, True, True)
    rng = np.random.default_rng(9)
    v, w = rng.standard_normal(3001), rng.standard_normal(50000)
    ctx.set_option('slab_width', slab)
    out = {}
    try:
        for part, sc, lpt in ((1, 0, 0), (2, 0, 0), (0, 0, 0), (0, 9, 0), (1, 0, 1), (2, 0, 1), (0, 0, 1), (0, 9, 1)):
            ctx.set_option('sell_partition', part)
            ctx.set_option('sell_slice_cost', sc)
            ctx.set_option('sell_lpt', lpt)
            D = Sparse(X, center_predictor=True, add_intercept=True, ctx=ctx)
            out[(part, sc, lpt)] = (D.dot(v), D.Tdot(w))
    finally:
        ctx.set_option('sell_partition', 0)
        ctx.set_option('sell_slice_cost', 0)
        ctx.set_option('sell_lpt', 1)
        ctx.set_option('slab_width', 0)
    base = out[(1, 0, 0)]
    assert relerr(base[0], O.dot(v)) < 1e-13 and relerr(base[1], O.Tdot(w)) < 1e-13
    for key, (d, t) in out.items():
        assert np.array_equal(d, base[0]) and np.array_equal(t, base[1]), key


@pytest.mark.parametrize('n,p,stream', [(200, 30, 1), (5000, 1300, 1), (1001, 2049, 1), (3, 5, 1), (777, 4097, 1), (301, 9001, 1),
                                        (200, 30, 0), (5000, 1300, 0)])
def test_dense_products_vs_oracle(ctx, n, p, stream):
    """stream = 1: the one-pass TMA streaming kernel (row groups in shared memory; odd p / ragged last group / p beyond
    one column per thread; p = 9001 exceeds what the kernel holds and takes the two-pass kernels), 0: two-pass kernels."""
    _, Dense = _designs()
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, p))
    ctx.set_option('dense_stream', stream)
    try:
        for center in (False, True):
            for icpt in (False, True):
                D = Dense(X.copy(), center_predictor=center, add_intercept=icpt, ctx=ctx)
                O = co.DesignOracle(X, center, icpt)
                v, w, wt = rng.standard_normal(D.shape[1]), rng.standard_normal(n), rng.random(n)
                assert relerr(D.dot(v), O.dot(v)) < 1e-12
                assert relerr(D.Tdot(w), O.Tdot(w)) < 1e-12
                assert relerr(D.compute_fisher_info(wt, diag_only=True), O.fisher_diag(wt)) < 1e-12
    finally:
        ctx.set_option('dense_stream', 1)


def test_csr_upload_and_csc_construction_are_bit_exact(ctx):
    """CSC on the device == scipy's tocsr().tocsc() byte for byte, including unsorted rows and duplicates."""
    Sparse, _ = _designs()
    rs = np.random.RandomState(4)
    cases = [random_sparse(5000, 300, 0.03, 1), random_sparse(40000, 1200, 0.01, 2, binary=True)]
    # unsorted column indices within rows + duplicate entries (scipy keeps both)
    n, p, nnz = 700, 50, 9000
    indptr = np.sort(np.concatenate(([0, nnz], rs.randint(0, nnz, n - 1)))).astype(np.int32)
    indices = rs.randint(0, p, nnz).astype(np.int32)
    data = rs.standard_normal(nnz)
    cases.append(sp.csr_matrix((data, indices, indptr), shape=(n, p)))
    for X in cases:
        D = Sparse(X, center_predictor=False, add_intercept=False, ctx=ctx, pattern_only=False)
        ip, ix, dv = D.export_csr()
        assert np.array_equal(ip, X.indptr) and np.array_equal(ix, X.indices) and np.array_equal(dv, X.data)
        C = X.tocsc()
        ip, ix, dv = D.export_csc()
        assert ip.dtype == C.indptr.dtype == np.int32
        assert np.array_equal(ip, C.indptr) and np.array_equal(ix, C.indices) and np.array_equal(dv, C.data)
        # and the products on the unsorted / duplicated matrix are still right
        v = rs.standard_normal(X.shape[1])
        assert relerr(D.dot(v), X @ v) < 1e-12


def test_edge_cases(ctx):
    Sparse, _ = _designs()
    # empty rows and ragged rows (every column keeps some variance, so none is dropped)
    rows = np.array([0, 0, 2, 5, 5, 5, 8, 8])
    cols = np.array([0, 3, 1, 0, 2, 3, 1, 2])
    X = sp.csr_matrix((np.arange(1., 9.), (rows, cols)), shape=(9, 4))
    for center in (False, True):
        D = Sparse(X, center_predictor=center, add_intercept=True, ctx=ctx)
        O = co.DesignOracle(X, center, True)
        v, w = np.arange(1., 6.), np.arange(1., 10.)
        assert np.allclose(D.dot(v), O.dot(v), rtol=1e-14, atol=1e-14)
        assert np.allclose(D.Tdot(w), O.Tdot(w), rtol=1e-14, atol=1e-14)
    with pytest.raises(ValueError):
        D.dot(np.ones(3))
    with pytest.raises(ValueError):
        D.Tdot(np.ones(3))
    full = D.compute_fisher_info(np.ones(9), diag_only=False)          # the full matrix (Cholesky comparator) on a tiny design
    A = O.toarray()
    assert np.allclose(full, A.T @ A, rtol=1e-13, atol=1e-13)
    # an all-zero matrix: every column is constant and is dropped (as the reference does) -> intercept only
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        Z = Sparse(sp.csr_matrix((6, 4)), center_predictor=True, add_intercept=True, ctx=ctx)
    assert Z.shape == (6, 1) and Z.nnz == 0
    assert np.array_equal(Z.dot(np.array([2.5])), np.full(6, 2.5))
    assert np.array_equal(Z.Tdot(np.arange(6.)), [15.0])
    assert np.array_equal(Z.compute_fisher_info(np.ones(6), diag_only=True), [6.0])


def test_constant_column_is_dropped_with_warning(ctx):
    Sparse, Dense = _designs()
    rng = np.random.default_rng(1)
    X = rng.standard_normal((50, 6))
    X[:, 2] = 1.0
    import warnings
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        D = Dense(X.copy(), add_intercept=True, ctx=ctx)
        S = Sparse(sp.csr_matrix(X), add_intercept=True, ctx=ctx)
    assert D.shape == S.shape == (50, 6)       # 5 predictors + intercept
    assert any('Intercept column' in str(r.message) for r in rec)


def test_memoised_dot_and_counters(ctx):
    Sparse, _ = _designs()
    X = random_sparse(400, 30, 0.2, 8)
    D = Sparse(X, center_predictor=True, add_intercept=True, ctx=ctx)
    v = np.linspace(-1, 1, 31)
    D.memoize_dot(True)
    a = D.dot(v); b = D.dot(v.copy())
    assert a is b and D.get_dot_count() == (1, 0)
    D.memoize_dot(False)
    D.Tdot(np.ones(400))
    assert D.n_matvec == 2
    D.reset_matvec_count()
    assert D.n_matvec == 0


def test_full_size_properties_c3(ctx):
    """BASELINE config 3 (OHDSI-scale: binary sparse 100k x 20k, ~0.5 %, bench.py's generator) -- too large for the
    numpy oracle to be the checker of every entry, so size-independent properties are asserted instead:
    linearity of dot / Tdot, adjointness <Xv, w> = <v, X'w>, agreement with scipy on a row / column sample,
    and fisher-diag == Tdot of the weights for a 0/1 matrix."""
    import bench
    Sparse, _ = _designs()
    n, p, dens = bench.WORKLOADS['C3']
    X, _ = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
    X = X[:, np.asarray(X.sum(axis=0)).ravel() > 0].tocsr()      # the class would drop the empty (constant) columns itself
    D = Sparse(X, center_predictor=True, add_intercept=True, ctx=ctx)
    assert D.is_binary and D.shape == (n, X.shape[1] + 1)
    P = D.shape[1]
    rng = np.random.default_rng(0)
    v1, v2, w1, w2 = rng.standard_normal(P), rng.standard_normal(P), rng.standard_normal(n), rng.standard_normal(n)
    a, b = 0.7, -1.9
    assert relerr(D.dot(a * v1 + b * v2), a * D.dot(v1) + b * D.dot(v2)) < 1e-13
    assert relerr(D.Tdot(a * w1 + b * w2), a * D.Tdot(w1) + b * D.Tdot(w2)) < 1e-13
    lhs, rhs = np.dot(D.dot(v1), w1), np.dot(v1, D.Tdot(w1))
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), np.linalg.norm(D.dot(v1)) * np.linalg.norm(w1))
    # spot check against scipy on the un-centred, intercept-free part
    c = np.asarray(X.mean(axis=0)).ravel()
    rows = rng.choice(n, 2000, replace=False)
    ref_rows = v1[0] + X[rows] @ v1[1:] - c @ v1[1:]
    assert relerr(D.dot(v1)[rows], ref_rows) < 1e-12
    cols = rng.choice(X.shape[1], 2000, replace=False)
    ref_cols = X[:, cols].T @ w1 - w1.sum() * c[cols]
    assert relerr(D.Tdot(w1)[1 + cols], ref_cols) < 1e-12
    # 0/1 entries: sum_i w_i x_ij^2 == sum_i w_i x_ij
    wt = rng.random(n)
    uncentred = Sparse(X, center_predictor=False, add_intercept=True, ctx=ctx)
    assert relerr(uncentred.compute_fisher_info(wt, diag_only=True), uncentred.Tdot(wt)) < 1e-13


def test_loglik_and_gradient_on_the_device(ctx):
    """bb_loglik_and_gradient (chain initialisation, reg_coef_sampler.py:281-327 -> logistic_model.py:49-55,
    linear_model.py:13-24): value and gradient vs the numpy formulas on the oracle design, sparse and dense."""
    import bayesbridge_b200 as bb
    Sparse, Dense = _designs()
    rng = np.random.default_rng(12)
    n, p = 6000, 90
    Xs = random_sparse(n, p, 0.05, seed=4, binary=True)
    Xd = rng.standard_normal((n, p))
    for X, Cls in ((Xs, Sparse), (Xd, Dense)):
        D = Cls(X.copy(), center_predictor=True, add_intercept=True, ctx=ctx)
        O = co.DesignOracle(X, True, True)
        beta = rng.standard_normal(p + 1) * 0.3
        eta = O.dot(beta)
        # logit
        n_trial = rng.integers(1, 4, n).astype(float)
        n_success = rng.binomial(n_trial.astype(int), 1 / (1 + np.exp(-eta))).astype(float)
        model = bb.RegressionModel((n_success, n_trial), D, family='logit')
        ll, grad = model.compute_loglik_and_gradient(beta)
        ll_ref = np.sum(n_success * eta - n_trial * np.logaddexp(0, eta))
        grad_ref = O.Tdot(n_success - n_trial / (1 + np.exp(-eta)))
        assert abs(ll - ll_ref) <= 1e-12 * abs(ll_ref) and relerr(grad, grad_ref) < 1e-12
        ll2, g2 = model.compute_loglik_and_gradient(beta, loglik_only=True)
        assert ll2 == ll and g2 is None
        # linear, on the same design: the outcome resident on the device handle is switched to this model's
        y = eta + rng.standard_normal(n)
        lin = bb.RegressionModel(y, D, family='linear')
        prec = 0.7
        ll, grad = lin.compute_loglik_and_gradient(beta, prec)
        ll_ref = n * np.log(prec) / 2 - prec * np.sum((y - eta) ** 2) / 2
        assert abs(ll - ll_ref) <= 1e-12 * abs(ll_ref) and relerr(grad, prec * O.Tdot(y - eta)) < 1e-12
        # and back: the first model re-pushes its outcome
        ll3, _ = model.compute_loglik_and_gradient(beta + 0.0, loglik_only=True)
        assert abs(ll3 - np.sum(n_success * eta - n_trial * np.logaddexp(0, eta))) <= 1e-12 * abs(ll3)


def test_empty_and_constant_columns_are_dropped(ctx):
    """remove_intercept_indicator (abstract_matrix.py:93-107) with the moments taken on the device: an empty column and
    a column of ones both go, with the reference's warning; the centring offsets equal the host means."""
    Sparse, _ = _designs()
    X = random_sparse(500, 12, 0.2, seed=1, hot_column=False).tolil()
    X[:, 3] = 0.0
    X[:, 7] = 1.0
    X = X.tocsr()
    X.eliminate_zeros()
    import warnings
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        D = Sparse(X, center_predictor=True, add_intercept=True, ctx=ctx)
    assert any('Intercept column' in str(r.message) for r in rec)
    keep = [j for j in range(12) if j not in (3, 7)]
    assert D.shape == (500, 11)
    assert np.allclose(D.column_offset, np.asarray(X[:, keep].mean(axis=0)).ravel(), rtol=1e-14, atol=0)
    O = co.DesignOracle(X[:, keep], True, True)
    v = np.random.default_rng(0).standard_normal(11)
    assert relerr(D.dot(v), O.dot(v)) < 1e-13
