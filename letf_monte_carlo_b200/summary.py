"""Post-Monte-Carlo statistics in the reference's definitions (letf/reporting.py:353-388, 320-325, 541-555),
computed from the engine's result arrays instead of per-trial Python dicts.

`summary_table` is the array-backed equivalent of the rows `create_summary_statistics` assembles before it
prints: exact `np.percentile` (linear interpolation) at P10/25/40/60/75/90 and the median of Final_Wealth,
`CAGR = (W / INITIAL_CAPITAL) ** (1 / years) - 1`, median Max_DD, mean Trades_Per_Year and the win rate
against S2 **by trial index**.  `histogram_quantile` reads the same percentiles off the allreduced
4096-bin histograms of `letf_reduce` (the multi-GPU summary; bin-width accuracy).
"""
from __future__ import annotations

from typing import Mapping

import numpy as np

from . import abi

PERCENTILES = (10, 25, 40, 50, 60, 75, 90)


def _arrays(results):
    """(final_wealth, max_dd, num_trades) from a StrategyResults view or a list of reference-style dicts."""
    if hasattr(results, "final_wealth"):
        return np.asarray(results.final_wealth), np.asarray(results.max_dd), np.asarray(results.num_trades), None
    fw = np.array([r["Final_Wealth"] for r in results], dtype=float)
    dd = np.array([r.get("Max_DD", 0) for r in results], dtype=float)
    tpy = np.array([r.get("Trades_Per_Year", 0) for r in results], dtype=float)
    return fw, dd, None, tpy


def summary_table(mc_results: Mapping, time_horizon: float, initial_capital: float = 10000.0,
                  benchmark_id: str = "S2", names: Mapping | None = None) -> list[dict]:
    """One row per strategy, sorted by median wealth (descending) like the reference's table."""
    rows = []
    spy = None
    if benchmark_id in mc_results and len(mc_results[benchmark_id]):
        spy = _arrays(mc_results[benchmark_id])[0]
    for sid, results in mc_results.items():
        if not len(results):
            continue
        fw, dd, nt, tpy = _arrays(results)
        pct = {f"p{q}": float(np.percentile(fw, q)) for q in PERCENTILES if q != 50}
        median = float(np.median(fw))
        win = float(np.mean(fw > spy) * 100) if spy is not None and len(spy) == len(fw) else 0.0
        trades = float(np.mean(nt) / time_horizon) if nt is not None else float(np.mean(tpy))
        rows.append({"id": sid, "name": (names or {}).get(sid, {}).get("name", sid) if names else sid,
                     **pct, "median": median, "cagr": (median / initial_capital) ** (1 / time_horizon) - 1,
                     "win": win, "dd": float(np.median(dd)), "trades": trades})
    rows.sort(key=lambda r: r["median"], reverse=True)
    return rows


def cagr(wealth, time_horizon: float, initial_capital: float = 10000.0):
    return (np.asarray(wealth, dtype=float) / initial_capital) ** (1 / time_horizon) - 1


def histogram_quantile(hist: np.ndarray, lo: float, hi: float, q: float) -> float:
    """q-th percentile (0..100) of a letf_reduce histogram (bins equally spaced on [lo, hi]); linear inside the bin."""
    hist = np.asarray(hist, dtype=np.float64)
    n = hist.sum()
    if n <= 0:
        return float("nan")
    target = q / 100.0 * n
    c = np.cumsum(hist)
    b = int(np.searchsorted(c, target, side="left"))
    b = min(b, len(hist) - 1)
    prev = c[b - 1] if b > 0 else 0.0
    frac = (target - prev) / hist[b] if hist[b] > 0 else 0.0
    width = (hi - lo) / len(hist)
    return lo + (b + frac) * width


def summary_from_histograms(hist, lo, hi, counters, initial_capital: float = 10000.0) -> list[dict]:
    """Per strategy: wealth percentiles from the log10(W/W0) histogram, median max drawdown, counters.
    hist: [S][3][bins] (metric 0 log10(W/W0), 1 CAGR, 2 max drawdown), counters: [S][4]."""
    out = []
    for s in range(hist.shape[0]):
        row = {f"p{q}": initial_capital * 10 ** histogram_quantile(hist[s, 0], lo[0], hi[0], q) for q in PERCENTILES}
        row["cagr_median"] = histogram_quantile(hist[s, 1], lo[1], hi[1], 50)
        row["dd_median"] = histogram_quantile(hist[s, 2], lo[2], hi[2], 50)
        valid = int(counters[s, 0])
        row.update(valid=valid, win=100.0 * int(counters[s, 1]) / max(valid, 1), total_trades=int(counters[s, 2]),
                   flagged=int(counters[s, 3]))
        out.append(row)
    return out


assert abi.N_METRICS == 3
