"""tests/golden/infer_regime_900d.npz: the LIVE reference's regime inference (letf/utils.py:55-127, as run_strategy_fixed
calls it, strategy.py:64-78) and its run_strategy_fixed results WITHOUT a caller regime path, on one seeded frame.
TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference)."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (the reference is imported only when this script runs as the generator: tests import `frame` on boxes without it)


def frame(T=900, seed=11):
    rng = np.random.default_rng(seed)
    vol = np.where((np.arange(T) // 150) % 2 == 0, 0.008, 0.024)
    spy = rng.normal(0.0003, vol)
    vix = np.clip(14 + 900 * vol + rng.normal(0, 2.0, T), 9, 80)
    vix[[40, 41, 300]] = np.nan
    df = pd.DataFrame({"SPY_Ret": spy, "QQQ_Ret": 1.2 * spy + rng.normal(0, 0.003, T), "TQQQ_Ret": 3.4 * spy - 0.0004,
                       "UPRO_Ret": 3.0 * spy - 0.0003, "SSO_Ret": 2.0 * spy - 0.0002, "TMF_Ret": -0.8 * spy + rng.normal(0, 0.006, T),
                       "TLT_Ret": -0.25 * spy + rng.normal(0, 0.004, T), "VIX": vix,
                       "TNX": 3.0 + rng.normal(0, 0.1, T).cumsum() * 0.1, "IRX": 2.5 + rng.normal(0, 0.1, T).cumsum() * 0.1})
    for a in ("SPY", "QQQ", "TQQQ", "UPRO", "SSO", "TMF", "TLT"):
        df[f"{a}_Price"] = 100 * (1 + df[f"{a}_Ret"]).cumprod()
    df["Cash_Ret"] = 0.04 / 252
    return df


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference")
    from letf import strategy as ref_strategy
    from letf.utils import compute_high_vol_probability, infer_regime_from_vix
    df = frame()
    rv = df["SPY_Ret"].rolling(20, min_periods=5).std().bfill().fillna(0) * np.sqrt(252)
    ts = (df["TNX"] - df["IRX"]).values
    out = {"p_high": compute_high_vol_probability(df["VIX"].values, realized_vol=rv.values, term_spread=ts),
           "regime": infer_regime_from_vix(df["VIX"].values, realized_vol=rv.values, term_spread=ts),
           "regime_no_ts": infer_regime_from_vix(df["VIX"].values, realized_vol=rv.values),
           "p_high_vix_only": compute_high_vol_probability(df["VIX"].values)}
    dfv = df.copy(); dfv["VIX"] = dfv["VIX"].ffill()        # run_strategy_fixed itself needs a gap-free VIX column
    for sid in ("S4", "S8", "S16"):
        curve, nt = ref_strategy.run_strategy_fixed(dfv, sid, None, {}, apply_costs=True)
        out[f"{sid}_curve"] = np.asarray(curve); out[f"{sid}_trades"] = nt
        curve, nt = ref_strategy.run_strategy_fixed(dfv, sid, np.zeros(17, dtype=int), {}, apply_costs=True)   # length mismatch
        out[f"{sid}_curve_mismatch"] = np.asarray(curve)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "infer_regime_900d.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
