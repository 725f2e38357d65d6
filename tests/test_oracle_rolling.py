"""The oracle's restatement of pandas' rolling kernels, checked against pandas itself
(the reference's strategy rules are written with pandas rolling/shift, strategy.py)."""
import numpy as np
import pandas as pd
import pytest

from oracle import oracle as orc


def _series(seed, n, kind):
    rng = np.random.default_rng(seed)
    if kind == "ret":
        return rng.standard_t(4, n) * 0.02
    if kind == "price":
        return 100 * np.cumprod(1 + rng.normal(3e-4, 0.012, n))
    if kind == "downside":
        r = rng.normal(0, 0.03, n)
        return np.where(r < 0, r, 0.0)
    if kind == "flat":
        x = np.full(n, 0.00125)
        x[n // 2:] = rng.normal(0, 0.01, n - n // 2)
        return x
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["ret", "price", "downside", "flat"])
@pytest.mark.parametrize("w,mp", [(3, 1), (5, 1), (10, 1), (20, 1), (20, 20), (60, 1), (126, 1), (200, 1), (200, 200), (14, 14)])
def test_rolling_mean_sum_std_bit_exact(kind, w, mp):
    x = _series(7 + w, 1500, kind)
    s = pd.Series(x)
    r = s.rolling(w, min_periods=mp)
    np.testing.assert_array_equal(orc.roll_mean(x, w, mp), r.mean().values)
    np.testing.assert_array_equal(orc.roll_sum(x, w, mp), r.sum().values)
    np.testing.assert_array_equal(orc.roll_std(x, w, mp), r.std().values)


@pytest.mark.parametrize("w", [3, 5, 20])
@pytest.mark.parametrize("pz", [0.5, 0.8, 0.9])
def test_rolling_std_constant_runs(w, pz):
    """series with long runs of identical values (what the downside-vol windows of S10/S15/S17/S19
    see after >= 20 consecutive non-negative returns): pandas recomputes the window from scratch
    when a removal collapses the variance; the oracle must follow bit-for-bit."""
    rng = np.random.default_rng(5 + w)
    r = rng.normal(0, 0.03, 20000) * 10 ** rng.uniform(-3, 0, 20000)
    r[rng.random(20000) < pz] = 0.01
    x = np.where(r < 0, r, 0.0)
    got = orc.roll_std(x, w, 1)
    ref = pd.Series(x).rolling(w, min_periods=1).std().values
    assert (ref[1:] == 0).sum() > 50            # the corner is exercised
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("kind", ["ret", "downside", "flat"])
def test_rolling_skew_close(kind):
    x = _series(11, 2000, kind)
    got = orc.roll_skew(x, 60, 1)
    ref = pd.Series(x).rolling(60, min_periods=1).skew().values
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    m = ~np.isnan(ref)
    np.testing.assert_allclose(got[m], ref[m], rtol=0, atol=5e-12)
