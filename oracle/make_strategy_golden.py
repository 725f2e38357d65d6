"""Golden for the run_strategy_fixed seam: the LIVE reference's `run_strategy_fixed` (letf/strategy.py:55) on one
seeded frame, equity curves and trade journals of a spread of rules.

TEST INFRASTRUCTURE, authoring container only (needs /root/reference).  The frame is rebuilt from the seed by
`strategy_frame()` wherever the fixture is used.

usage: python oracle/make_strategy_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
SIDS = ["S1", "S2", "S3", "S4", "S5", "S6", "S7", "S8", "S10", "S13", "S15", "S17", "S19"]
T, SEED = 700, 4711


def strategy_frame(T=T, seed=SEED):
    """(df, regime_path): the columns engine.py:537-550 assembles - <asset>_Ret, Cash_Ret, <asset>_Price, TLT_Ret,
    TLT_Price, VIX - from seeded fat-tailed returns with two volatility states."""
    import pandas as pd
    from letf_monte_carlo_b200 import abi, config as cfg
    rng = np.random.default_rng(seed)
    regime = np.zeros(T, dtype=int)
    t, cur = 0, 0
    while t < T:
        ln = int(rng.geometric(1 / (90 if cur == 0 else 35)))
        regime[t:t + ln] = cur
        t += ln
        cur ^= 1
    vol = np.where(regime == 0, 0.009, 0.024)
    spy = 0.0004 + vol * rng.standard_t(5, T) * 0.75
    qqq = 1.15 * spy + 0.005 * rng.standard_normal(T)
    tlt = -0.15 * spy + 0.005 * rng.standard_normal(T)
    lev = {"TQQQ": (3.0, qqq), "UPRO": (3.0, spy), "SSO": (2.0, spy), "TMF": (3.0, tlt), "SPY": (1.0, spy), "QQQ": (1.0, qqq)}
    cols = {}
    for a in abi.ASSET_NAMES:
        L, u = lev[a]
        cols[f"{a}_Ret"] = np.clip(L * u - 0.00004 * (L > 1) + 0.0003 * rng.standard_normal(T), -0.999, 10.0)
    df = pd.DataFrame(cols)
    df["Cash_Ret"] = np.where(regime == 0, cfg.CASH_RATE_BY_REGIME[0], cfg.CASH_RATE_BY_REGIME[1]) / 252
    for a in abi.ASSET_NAMES:
        df[f"{a}_Price"] = (1 + df[f"{a}_Ret"].fillna(0)).cumprod() * 100
    df["TLT_Ret"] = tlt
    df["TLT_Price"] = (1 + df["TLT_Ret"].fillna(0)).cumprod() * 100
    df["VIX"] = np.where(regime == 0, 16.0, 31.0) + 4.0 * rng.standard_normal(T)
    return df, regime


def main():
    sys.path.insert(0, REF)
    from letf import strategy as ref_strategy
    from letf.trade import TradeJournal, TAXABLE_IDS
    df, regime = strategy_frame()
    out = {"sids": np.array(SIDS), "T": T, "seed": SEED}
    for sid in SIDS:
        j = TradeJournal() if sid in TAXABLE_IDS else None
        ec, nt = ref_strategy.run_strategy_fixed(df, sid, regime, {}, apply_costs=True, trade_journal=j)
        out[f"{sid}_curve"] = np.asarray(ec.values, dtype=float)
        out[f"{sid}_trades"] = int(nt)
        tl = j.get_full_trades() if j is not None else []
        out[f"{sid}_journal_n"] = len(tl)
        out[f"{sid}_journal_volume"] = float(sum(t["dollar_amount"] for t in tl))
        out[f"{sid}_journal_days"] = np.array([t["day_index"] for t in tl], dtype=np.int32)
    import pandas, scipy
    out["versions"] = np.array([f"numpy {np.__version__}", f"pandas {pandas.__version__}", f"scipy {scipy.__version__}"])
    out["generator"] = np.array("oracle/make_strategy_golden.py: live reference letf.strategy.run_strategy_fixed on strategy_frame()")
    path = os.path.join(ROOT, "tests", "golden", "run_strategy_fixed_700d.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {s: int(out[f"{s}_trades"]) for s in SIDS})


if __name__ == "__main__":
    main()
