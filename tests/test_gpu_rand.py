"""GPU: Polya-Gamma / tilted-stable / Gaussian device kernels.
* near-exact: the oracle port driven by the device's own Philox streams must give the same variates;
* distributional: two-sample KS against samples of the compiled reference (golden/ks_ref.npz) and analytic moments;
* determinism under (seed, offset) and invariance to how observations are sharded."""
import numpy as np
import pytest
from scipy.stats import ks_2samp

from conftest import golden, import_reference, record_achieved
from oracle import philox_port as pp

pytestmark = pytest.mark.gpu


def _samplers(ctx, seed):
    from bayesbridge_b200.random import DevicePolyaGamma, DeviceTiltedStable
    return DevicePolyaGamma(ctx, seed), DeviceTiltedStable(ctx, seed)


def test_philox_normals_match_the_port(ctx):
    from bayesbridge_b200 import _lib
    out = np.empty(300)
    _lib.check(_lib.load().bb_philox_normal(ctx.handle, 300, 1, (5 << 32) + 77, (1 << 33) + 9, (1 << 32) + 5, _lib.dptr(out)))
    ref = pp.philox_normals(300, 1, (5 << 32) + 77, (1 << 33) + 9, (1 << 32) + 5)
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-14)


def test_pg_matches_port_on_device_streams(ctx):
    pg, _ = _samplers(ctx, 31)
    rng = np.random.default_rng(2)
    shape = rng.integers(1, 4, 400).astype(np.int32)
    tilt = np.concatenate((rng.standard_normal(200) * 2, rng.standard_normal(200) * 25))
    tilt[:3] = (0.0, 1e-9, 3.14159)
    got = pg.rand_polyagamma(shape, tilt, index_offset=1000)
    ref = pp.pg_with_device_streams(shape, tilt, 31, 0, 1000)
    close = np.isclose(got, ref, rtol=1e-10, atol=0)
    assert close.mean() > 0.99, "accept/reject decisions may flip on ulp differences, but only rarely"
    assert np.allclose(got[close], ref[close], rtol=1e-10)


def test_ts_matches_port_on_device_streams(ctx):
    _, ts = _samplers(ctx, 8)
    rng = np.random.default_rng(3)
    for a in (1 / 32, 0.25, 0.5):
        tilt = np.exp(rng.standard_normal(150) * 4)
        got = ts.sample(a, tilt)
        ref = pp.ts_with_device_streams(a, tilt, 8, ts.offset - 1)
        close = np.isclose(got, ref, rtol=1e-8, atol=0)
        assert close.mean() > 0.97


def test_pg_distribution_vs_reference_samples_and_moments(ctx):
    pg, _ = _samplers(ctx, 5)
    g = golden('ks_ref.npz')
    N = 200000
    for (b, c), ref in zip(g['pg_grid'], g['pg_samples']):
        x = pg.rand_polyagamma(np.full(N, int(b), dtype=np.int32), np.full(N, c))
        mean = b / 4 if c < 1e-5 else b / (2 * c) * np.tanh(c / 2)
        var = b / 24 if c < 1e-5 else b * (np.sinh(min(c, 300.)) - min(c, 300.)) / (4 * min(c, 300.) ** 3 * np.cosh(min(c, 300.) / 2) ** 2)
        assert abs(x.mean() - mean) < 5 * np.sqrt(var / N), (b, c)
        assert abs(x.var() / var - 1) < 0.05, (b, c)
        assert ks_2samp(x, ref).pvalue > 1e-4, (b, c)


def test_ts_distribution_vs_reference_samples(ctx):
    _, ts = _samplers(ctx, 6)
    g = golden('ks_ref.npz')
    N = 100000
    for (a, t), ref in zip(g['ts_grid'], g['ts_samples']):
        x = ts.sample(float(a), np.full(N, t))
        assert np.all(np.isfinite(x)) and np.all(x > 0)
        assert ks_2samp(x, ref).pvalue > 1e-4, (a, t)
        # Laplace transform identity of the tilted stable law: E[exp(-s X)] = exp(t^a - (t+s)^a)
        s = t
        assert np.mean(np.exp(-s * x)) == pytest.approx(np.exp(t ** a - (t + s) ** a), rel=0.02)


def _ks_distance_to_table(x, quantiles):
    """sup |F_x - F_ref| with F_ref interpolated from the reference's quantile table (golden_ks_quantiles)."""
    probs = np.linspace(0.0, 1.0, len(quantiles))
    x = np.sort(x)
    n = len(x)
    F = np.interp(x, quantiles, probs)
    i = np.arange(n)
    return max(np.abs((i + 1) / n - F).max(), np.abs(i / n - F).max())


# two-sample KS critical distance at n = m = 1e6, alpha = 1e-4: sqrt(-ln(alpha/2)/2) * sqrt(2/n) = 3.15e-3, plus the
# resolution of the 4001-point quantile table (2.5e-4).  The reference against its own table gives 0.6e-3 ... 2.4e-3.
KS_BOUND_1E6 = 3.4e-3


def test_pg_and_ts_ks_at_one_million_draws(ctx):
    """N = 1e6 draws per grid cell against 1e6 draws of the compiled reference sampler
    (summarised as quantile tables, tests/golden/make_golden.py::golden_ks_quantiles)."""
    pg, ts = _samplers(ctx, 77)
    g = golden('ks_quantiles_ref.npz')
    N = 1_000_000
    assert int(g['n_reference_draws']) == N
    for (b, c), q in zip(g['pg_grid'], g['pg_quantiles']):
        d = _ks_distance_to_table(pg.rand_polyagamma(np.full(N, int(b), dtype=np.int32), np.full(N, c)), q)
        record_achieved('ks_1e6_polya_gamma', (int(b), float(c)), d, KS_BOUND_1E6)
        assert d < KS_BOUND_1E6, ('pg', b, c, d)
    for (a, t), q in zip(g['ts_grid'], g['ts_quantiles']):
        d = _ks_distance_to_table(ts.sample(float(a), np.full(N, t)), q)
        record_achieved('ks_1e6_tilted_stable', (float(a), float(t)), d, KS_BOUND_1E6)
        assert d < KS_BOUND_1E6, ('ts', a, t, d)


def test_pg_vs_live_reference_large_sample(ctx):
    if import_reference() is None:
        pytest.skip('oracle/_ref not built')
    from bayesbridge.random.polya_gamma import PolyaGammaDist
    pg, _ = _samplers(ctx, 7)
    N = 500000
    for b, c in ((1, 0.01), (2, 100.), (1, 2.0)):      # the reference notebook's two points + one
        x = pg.rand_polyagamma(np.full(N, b, dtype=np.int32), np.full(N, c))
        r = PolyaGammaDist(9).rand_polyagamma(np.full(N, b, dtype=np.intc), np.full(N, c))
        assert ks_2samp(x, r).pvalue > 1e-4


def test_determinism_offsets_and_sharding_invariance(ctx):
    from bayesbridge_b200.random import DevicePolyaGamma
    shape = np.ones(1000, dtype=np.int32)
    tilt = np.linspace(-5, 5, 1000)
    a = DevicePolyaGamma(ctx, 42).rand_polyagamma(shape, tilt)
    b = DevicePolyaGamma(ctx, 42).rand_polyagamma(shape, tilt)
    assert np.array_equal(a, b)
    g = DevicePolyaGamma(ctx, 42)
    first, second = g.rand_polyagamma(shape, tilt), g.rand_polyagamma(shape, tilt)
    assert np.array_equal(first, a) and not np.array_equal(second, a)       # offset advances per call
    state = g.get_state()
    third = g.rand_polyagamma(shape, tilt)
    g2 = DevicePolyaGamma(ctx, 1); g2.set_state(state)
    assert np.array_equal(g2.rand_polyagamma(shape, tilt), third)          # resumable from (seed, offset)
    # two shards of 500 with global index offsets == one call of 1000
    h = DevicePolyaGamma(ctx, 42)
    lo = h.rand_polyagamma(shape[:500], tilt[:500], index_offset=0)
    h.offset = 0
    hi = h.rand_polyagamma(shape[500:], tilt[500:], index_offset=500)
    assert np.array_equal(np.concatenate((lo, hi)), a)
    # independence across observations: adjacent draws with identical parameters are uncorrelated
    x = DevicePolyaGamma(ctx, 3).rand_polyagamma(np.ones(200000, dtype=np.int32), np.zeros(200000))
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 0.01


def test_input_validation(ctx):
    pg, ts = _samplers(ctx, 1)
    with pytest.raises(TypeError):
        pg.rand_polyagamma([1, 2], np.zeros(2))
    with pytest.raises(ValueError):
        pg.rand_polyagamma(np.ones(3, dtype=np.int32), np.zeros(2))
    with pytest.raises(ValueError):
        pg.rand_polyagamma(np.ones(2), np.zeros(2))
    with pytest.raises(ValueError):
        ts.sample(0.5, np.array([1.0, -1.0]))
    with pytest.raises(ValueError):
        ts.sample(1.5, np.array([1.0]))
    assert pg.rand_polyagamma(np.zeros(0, dtype=np.int32), np.zeros(0)).size == 0
