"""Sampled-trial tax path of the taxable strategies (reference: letf/reporting.py:399-560).

What the reference does per taxable strategy: rank the trials by pre-tax wealth, take <= 50 evenly spaced
ranks (`MAX_TAX_SIMS`), push each sampled trial's `Trade_List` through the tax engine
(`letf.integration.process_trades_with_wired_engine`), interpolate the tax of every other trial on its pre-tax
wealth, and report percentiles of both distributions, tax drag and the post-tax win rate against S2.

On the GPU path only the sampled trials ever have a trade list: it is replayed on the device for exactly those
trials (`letf_trade_lists_host`, one launch per strategy), everything else stays as arrays.  The tax engine (lot
selection, wash sales, yearly netting, marginal brackets: host bookkeeping on <= 50 trade lists per strategy) is
`tax_engine.process_trades`, a restatement of the reference's `process_trades_with_wired_engine` pinned to it
(tests/test_tax_engine.py); any other engine can be injected as `tax_fn`.
"""
from __future__ import annotations

from typing import Callable, Mapping, Sequence

import numpy as np

from .trade import TAXABLE_IDS

MAX_TAX_SIMS = 50                                      # reporting.py:455
PCTS = (("p10", 10), ("p25", 25), ("p40", 40), ("p60", 60), ("p75", 75), ("p90", 90))   # reporting.py:513


def default_tax_fn() -> Callable:
    """The in-tree host tax engine (tax_engine.py: wash sales, lot selection, yearly netting, marginal brackets - the
    reference's `process_trades_with_wired_engine`, letf/integration.py:26, restated and pinned to it) with the reference's
    default elections (FIFO lots, short-term-first loss ordering)."""
    from .tax_engine import process_trades

    def fn(trades, time_horizon_years, initial_capital, strategy_id, tax_config):
        return process_trades(trades, time_horizon_years, initial_capital, tax_config=tax_config)["total_tax"]
    return fn


def tax_sample_indices(wealth: np.ndarray, max_tax_sims: int = MAX_TAX_SIMS) -> list[int]:
    """reporting.py:457-467: every trial when n <= 50, else the trials at evenly spaced wealth ranks."""
    n = len(wealth)
    if n <= max_tax_sims:
        return list(range(n))
    order = np.argsort(wealth)
    positions = np.linspace(0, n - 1, max_tax_sims, dtype=int)
    return [int(order[p]) for p in positions]


def _wealth_of(results) -> np.ndarray:
    if hasattr(results, "final_wealth"):
        return np.asarray(results.final_wealth, dtype=float)
    return np.array([r["Final_Wealth"] for r in results], dtype=float)


def _trade_list(results, i):
    if hasattr(results, "trade_list"):
        return results.trade_list(i)                   # device replay of ONE trial (letf_trade_list_host)
    return results[i].get("Trade_List", [])


def interpolated_tax(results, sid: str, time_horizon, tax_fn: Callable, tax_config: Mapping | None,
                     initial_capital: float = 10000.0, max_tax_sims: int = MAX_TAX_SIMS, lists=None):
    """-> (tax of every trial, sampled trial indices, their taxes); reporting.py:469-506.
    lists: the sampled trials' Trade_Lists when the caller already replayed them (one launch for all strategies)"""
    wealth = _wealth_of(results)
    n = len(wealth)
    idx = tax_sample_indices(wealth, max_tax_sims)
    taxes = []
    if lists is None:
        lists = results.trade_lists(idx) if hasattr(results, "trade_lists") else [_trade_list(results, i) for i in idx]
    for i, tl in zip(idx, lists):
        taxes.append(float(tax_fn(tl, time_horizon, initial_capital, f"{sid}_bulk", tax_config)) if tl else 0.0)
    sw, st = wealth[idx], np.array(taxes)
    if n <= max_tax_sims:
        all_tax = np.zeros(n)
        all_tax[idx] = st
    else:
        o = np.argsort(sw)
        all_tax = np.interp(wealth, sw[o], st[o])
    return all_tax, idx, st


def taxable_summary(mc_results: Mapping, time_horizon, tax_fn: Callable | None = None, tax_config: Mapping | None = None,
                    initial_capital: float = 10000.0, taxable_ids: Sequence[str] = TAXABLE_IDS,
                    names: Mapping | None = None, benchmark_id: str = "S2", max_tax_sims: int = MAX_TAX_SIMS) -> list[dict]:
    """The `data` rows of the reference's taxable table (reporting.py:408-560), sorted by P60 post-tax wealth."""
    tax_fn = tax_fn or default_tax_fn()
    spy_post = None
    if benchmark_id in mc_results and len(mc_results[benchmark_id]):
        # reporting.py:416-438: S2's own tax, trial by trial (a buy & hold benchmark has no Trade_List: zero tax)
        res = mc_results[benchmark_id]
        spy_pre = _wealth_of(res)
        spy_post = spy_pre.copy()
        if benchmark_id in taxable_ids or not hasattr(res, "trade_list"):
            for i in range(len(spy_pre)):
                tl = _trade_list(res, i)
                if tl:
                    spy_post[i] = spy_pre[i] - float(tax_fn(tl, time_horizon, initial_capital, "SPY_tax", tax_config))
    rows = []
    # device-backed results: the sampled trials of ALL taxable strategies are replayed in one launch
    replayed = {}
    if hasattr(mc_results, "trade_lists_multi"):
        wanted = {sid: tax_sample_indices(_wealth_of(mc_results[sid]), max_tax_sims)
                  for sid in taxable_ids if sid in mc_results and len(mc_results[sid])}
        replayed = mc_results.trade_lists_multi(wanted) if wanted else {}
    for sid in taxable_ids:
        if sid not in mc_results or not len(mc_results[sid]):
            continue
        results = mc_results[sid]
        wealth = _wealth_of(results)
        all_tax, idx, st = interpolated_tax(results, sid, time_horizon, tax_fn, tax_config, initial_capital, max_tax_sims,
                                            lists=replayed.get(sid))
        post = wealth - all_tax
        pcts = {}
        for pname, pval in PCTS:
            pre_w, post_w = float(np.percentile(wealth, pval)), float(np.percentile(post, pval))
            pre_c = (pre_w / initial_capital) ** (1 / time_horizon) - 1
            post_c = (post_w / initial_capital) ** (1 / time_horizon) - 1 if post_w > 0 else float("nan")
            if pre_c > 0.001 and not np.isnan(post_c):
                drag = min((pre_c - post_c) / pre_c * 100, 100.0)
            elif pre_c > 0.001:
                drag = 100.0
            else:
                drag = 0
            pcts[pname] = {"pre_cagr": pre_c, "post_cagr": post_c, "drag": drag, "pre_wealth": pre_w,
                           "post_wealth": post_w, "tax_paid": pre_w - post_w}
        if hasattr(results, "max_dd"):
            max_dd = float(np.median(results.max_dd))
            trades = float(np.mean(results.num_trades) / time_horizon)
        else:
            max_dd = float(np.median([r.get("Max_DD", 0) for r in results]))
            trades = float(np.mean([r.get("Trades_Per_Year", 0) for r in results]))
        if spy_post is not None and len(spy_post) == len(wealth):
            win = float(np.sum(post > spy_post) / len(post) * 100)          # post-tax vs post-tax, by trial index
        else:
            win = 0
        rows.append({"id": sid, "name": (names or {}).get(sid, {}).get("name", sid), "pcts": pcts, "max_dd": max_dd,
                     "trades": trades, "win": win, "sampled_trials": idx, "sampled_tax": st})
    rows.sort(key=lambda r: r["pcts"]["p60"]["post_wealth"], reverse=True)
    return rows
