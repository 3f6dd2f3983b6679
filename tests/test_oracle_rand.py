"""CPU: the oracle's random-variate ports reproduce the compiled reference draw for draw."""
import numpy as np
import pytest

from conftest import golden, import_reference
from oracle.rand_port import PolyaGammaPort, TiltedStablePort, log_ndtr


def test_pg_port_bit_exact_vs_reference_fixture():
    g = golden('random_ref.npz')
    out = PolyaGammaPort(int(g['pg_seed'])).rand_polyagamma(g['pg_shape'], g['pg_tilt'])
    assert np.array_equal(out, g['pg_out'])


def test_ts_port_bit_exact_vs_reference_fixture():
    g = golden('random_ref.npz')
    for k in range(3):
        out = TiltedStablePort(int(g['ts_seed'])).sample(float(g['ts%d_char_exp' % k]), g['ts%d_tilt' % k])
        assert np.array_equal(out, g['ts%d_out' % k])


def test_log_ndtr_regimes():
    from scipy.special import log_ndtr as sp_log_ndtr
    for a in (-45.0, -25.0, -20.0, -19.9, -5.0, -1.25, 0.0, 3.0, 6.0, 6.1, 12.0):
        assert log_ndtr(a) == pytest.approx(float(sp_log_ndtr(a)), rel=1e-12, abs=3e-16)  # log(Phi) near 0 loses digits by construction


def test_pg_port_moments():
    b, c = 2, 1.5
    x = PolyaGammaPort(3).rand_polyagamma(np.full(4000, b), np.full(4000, c))
    mean = b / (2 * c) * np.tanh(c / 2)
    var = b * (np.sinh(c) - c) / (4 * c ** 3 * np.cosh(c / 2) ** 2)
    assert abs(x.mean() - mean) < 5 * np.sqrt(var / x.size)


def test_ports_vs_live_reference_when_built():
    ref = import_reference()
    if ref is None:
        pytest.skip('oracle/_ref not built')
    from bayesbridge.random.polya_gamma import PolyaGammaDist
    rng = np.random.default_rng(21)
    shape = rng.integers(1, 4, 60).astype(np.intc)
    tilt = rng.standard_normal(60) * 8
    assert np.array_equal(PolyaGammaDist(77).rand_polyagamma(shape, tilt),
                          PolyaGammaPort(77).rand_polyagamma(shape, tilt))
