"""tests/golden/tax/engine_cases.json: the LIVE reference's tax engine (letf/integration.py:26-556 over letf/tax/*) on a
synthetic stress set of trade lists - several assets, partial sells, loss sales re-bought inside and outside the 30-day
window, same-day trades, multi-year holding - under every loss-carry-forward election, every lot-selection method, both
filing statuses, several states and a retirement scenario.  Stored: the trade lists and what the reference returned
(total tax, per-year taxable amounts / tax / carry-forwards).  TEST INFRASTRUCTURE; authoring container only.

usage: python oracle/make_tax_engine_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOSS = ["offset_st_first", "offset_lt_first", "defer_maximum", "use_all_asap"]
LOTS = ["fifo", "lifo", "hifo", "lofo", "ltfo", "stfo", "mintax", "specific"]
CONFIGS = [
    {"filing_status": "single", "state": "CA", "ordinary_income": 150000, "career_stage": "mid"},
    {"filing_status": "married", "state": "NY", "ordinary_income": 320000},
    {"filing_status": "single", "state": "WA", "ordinary_income": 90000},
    {"filing_status": "married", "state": "TX", "ordinary_income": 60000, "years_until_retirement": 3},
    {"filing_status": "single", "state": "NJ", "ordinary_income": 210000, "years_until_retirement": 2, "retirement_income": 55000},
    {"filing_status": "single", "state": "IL", "ordinary_income": 40000},
]


def trade_list(seed, years, n_assets, n_trades):
    """a random but plausible journal: alternating accumulation / partial liquidation per asset on a drifting price path"""
    rng = np.random.default_rng(seed)
    assets = ["TQQQ", "SPY", "TMF", "UPRO"][:n_assets]
    T = years * 252
    px = {a: 100 * np.exp(np.cumsum(rng.normal(0.0002, 0.03 if a != "SPY" else 0.011, T))) for a in assets}
    held = {a: 0.0 for a in assets}
    days = np.sort(rng.integers(0, T, n_trades))
    if seed % 3 == 0:                       # clusters of trades a few days apart: wash-sale windows
        days = np.sort(np.concatenate([days[: n_trades // 2], days[: n_trades // 2] + rng.integers(1, 25, n_trades // 2)])) % T
        days.sort()
    out = []
    for d in days:
        a = assets[int(rng.integers(0, n_assets))]
        p = float(px[a][d])
        if held[a] > 0.5 and rng.random() < 0.5:
            sh = float(held[a] * rng.choice([0.25, 0.5, 1.0]))
            act = "SELL"
            held[a] -= sh
        else:
            sh = float(rng.uniform(5, 60))
            act = "BUY"
            held[a] += sh
        t = {"day_index": int(d), "asset": a, "action": act, "shares": sh, "price": p, "dollar_amount": sh * p}
        if seed % 4 == 1 and rng.random() < 0.3:
            del t["shares"]                 # the engines fall back to dollar_amount / price
        out.append(t)
    return out


def cases():
    cs = []
    for k in range(48):
        years = [2, 3, 5, 10][k % 4]
        cs.append({"seed": 1000 + k, "years": years, "n_assets": 1 + k % 4, "n_trades": [12, 40, 90, 400][(k // 4) % 4],
                   "loss": LOSS[k % 4], "lots": LOTS[k % 8], "config": CONFIGS[k % len(CONFIGS)], "capital": [10000.0, 250000.0][k % 2]})
    # the benchmarked horizon: 30 years of near-daily trading (what S7 ... S19 journal), thousands of buys per asset - the
    # shapes on which a per-sale walk over all buys is quadratic (the in-tree engine bisects running sums instead)
    for k in range(4):
        cs.append({"seed": 2000 + k, "years": 30, "n_assets": 1 + k % 2, "n_trades": [3000, 5000, 6000, 4000][k],
                   "loss": LOSS[k % 4], "lots": ["fifo", "fifo", "hifo", "lifo"][k], "config": CONFIGS[k % len(CONFIGS)], "capital": 10000.0})
    return cs


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference")
    from letf.integration import process_trades_with_wired_engine
    from letf.tax.engine import CapitalLossUsageStrategy, LotSelectionMethod, TaxpayerElections
    docs = []
    for c in cases():
        tl = trade_list(c["seed"], c["years"], c["n_assets"], c["n_trades"])
        el = TaxpayerElections(capital_loss_strategy=CapitalLossUsageStrategy(c["loss"]), lot_selection_method=LotSelectionMethod(c["lots"]))
        r = process_trades_with_wired_engine(trades=[dict(t) for t in tl], time_horizon_years=c["years"], elections=el,
                                             initial_capital=c["capital"], debug=False, strategy_id="X", tax_config=dict(c["config"]))
        docs.append({"case": c, "total_tax": r["total_tax"], "final_cf": r["final_cf"],
                     "yearly": [{k: float(y[k]) for k in ("taxable_st", "taxable_lt", "tax", "st_cf", "lt_cf", "capital_loss_deduction",
                                                          "federal_total", "state_tax", "niit_tax")} for y in r["yearly_taxes"]],
                     "income": [float(v) for v in r["income_trajectory"]["income"]]})
    path = os.path.join(ROOT, "tests", "golden", "tax", "engine_cases.json")
    with open(path, "w") as f:
        json.dump({"generator": "oracle/make_tax_engine_golden.py: live reference process_trades_with_wired_engine", "cases": docs}, f)
    print(len(docs), "cases,", sum(d["total_tax"] > 0 for d in docs), "with tax > 0, file", os.path.getsize(path), "bytes;",
          "max total tax %.0f" % max(d["total_tax"] for d in docs))
