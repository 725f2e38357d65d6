"""Latency of the single-frame `run_strategy_fixed` drop-in (what letf/historical.py calls once per strategy on the historical
frame): all 19 strategies on one 7 000-day frame, first and second pass.  usage: python scripts/run_strategy_latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from letf_monte_carlo_b200 import run_strategy_fixed, synthetic
from letf_monte_carlo_b200.trade import TradeJournal

T = 7000
rng = np.random.default_rng(1)
ns = synthetic.config_namespace()
cols = {}
for a, s in zip(("TQQQ", "UPRO", "SSO", "TMF", "SPY", "QQQ"), (0.03, 0.03, 0.02, 0.025, 0.01, 0.013)):
    r = rng.standard_t(5, T) * s + 0.0003
    cols[f"{a}_Ret"] = r
    cols[f"{a}_Price"] = 100 * np.cumprod(1 + r)
cols["TLT_Ret"] = rng.normal(0, 0.008, T)
cols["VIX"] = 15 + 10 * rng.random(T)
cols["IRX"] = np.full(T, 4.5)
cols["Cash_Ret"] = np.full(T, 0.045 / 252)
df = pd.DataFrame(cols, index=pd.bdate_range("1995-01-02", periods=T))
regime = (rng.random(T) < 0.2).astype(int)
for label in ("first pass", "second pass"):
    t_all = time.perf_counter()
    worst = 0.0
    for sid in ns.STRATEGIES:
        t0 = time.perf_counter()
        eq, nt = run_strategy_fixed(df, sid, regime, None, True, TradeJournal(), cfg=ns)
        worst = max(worst, time.perf_counter() - t0)
    print(f"{label}: 19 strategies {1e3 * (time.perf_counter() - t_all):8.1f} ms (slowest call {1e3 * worst:.1f} ms), last equity {float(eq.iloc[-1]):.2f}")
