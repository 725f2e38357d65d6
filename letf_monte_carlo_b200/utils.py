"""Host-side frame preparation of the `run_strategy_fixed` seam.

The reference infers a regime path from the frame when the caller passes none, or one whose length does not match the
frame (letf/strategy.py:64-78; this is how letf/historical.py:96 calls it): a logistic stress score on VIX, 20-day realised
SPY volatility and the term spread, EWMA-smoothed, then a two-threshold hysteresis switch (letf/utils.py:55-127).  It is
O(T) scalar work on one frame, done here on the host before the frame is uploaded.
"""
from __future__ import annotations

import numpy as np


def _ffill_bfill(x: np.ndarray, default: float) -> np.ndarray:
    """pandas `.ffill().bfill().fillna(default)` on a float vector"""
    x = np.asarray(x, dtype=float).copy()
    ok = ~np.isnan(x)
    if not ok.any():
        x[:] = default
        return x
    idx = np.where(ok, np.arange(len(x)), -1)
    np.maximum.accumulate(idx, out=idx)
    first = int(np.argmax(ok))
    idx[idx < 0] = first
    return x[idx]


def compute_high_vol_probability(vix_series, realized_vol=None, term_spread=None, smoothing: float = 0.94) -> np.ndarray:
    """P(high-volatility regime) per day in [0.001, 0.999] (letf/utils.py:55-93)."""
    vix = np.asarray(vix_series, dtype=float)
    n = len(vix)
    if realized_vol is None:
        import pandas as pd
        rv = pd.Series(vix).rolling(20, min_periods=5).std().bfill().fillna(0).to_numpy() / 100.0
    else:
        rv = np.asarray(realized_vol, dtype=float)
        rv = _ffill_bfill(rv, np.nanmedian(rv) if np.isfinite(rv).any() else np.nan)
    ts = np.zeros(n) if term_spread is None else _ffill_bfill(np.asarray(term_spread, dtype=float), 0.0)
    logit = -4.0 + 0.22 * (np.where(np.isnan(vix), 20.0, vix) - 20.0) + 6.5 * (np.where(np.isnan(rv), 0.18, rv) - 0.18) \
        + 0.10 * np.clip(-ts, -5, 5)
    raw = 1.0 / (1.0 + np.exp(-np.clip(logit, -20, 20)))
    p = np.zeros(n)
    if n:
        p[0] = raw[0]
    for i in range(1, n):                       # EWMA recursion, same operation order as the reference loop
        p[i] = smoothing * p[i - 1] + (1 - smoothing) * raw[i]
    return np.clip(p, 0.001, 0.999)


def infer_regime_from_vix(vix_series, realized_vol=None, term_spread=None, hysteresis: float = 0.08) -> np.ndarray:
    """Regime path (0 = low vol, 1 = high vol) by hysteresis around P = 0.5 (letf/utils.py:96-127)."""
    p = compute_high_vol_probability(vix_series, realized_vol=realized_vol, term_spread=term_spread)
    up, down = 0.50 + hysteresis / 2, 0.50 - hysteresis / 2
    out = np.zeros(len(p), dtype=int)
    if not len(p):
        return out
    state = 1 if p[0] >= 0.50 else 0
    out[0] = state
    for i in range(1, len(p)):
        if state == 0 and p[i] >= up:
            state = 1
        elif state == 1 and p[i] <= down:
            state = 0
        out[i] = state
    return out


def regime_path_for_frame(df, regime_path):
    """The regime path `run_strategy_fixed` uses for `df` (strategy.py:64-78): the caller's when its length matches the
    frame, else inferred from VIX / realised volatility / term spread, else all zeros."""
    T = len(df)
    if regime_path is not None and len(regime_path) == T:
        return np.asarray(regime_path, dtype=np.int64)
    if "VIX" in df.columns:
        rv = df["SPY_Ret"].rolling(20, min_periods=5).std().bfill().fillna(0) * np.sqrt(252)
        ts = (df["TNX"] - df["IRX"]).to_numpy() if "TNX" in df.columns and "IRX" in df.columns else None
        return infer_regime_from_vix(df["VIX"].to_numpy(), realized_vol=rv.to_numpy(), term_spread=ts)
    return np.zeros(T, dtype=np.int64)
