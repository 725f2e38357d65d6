"""Latency of the reference's DEFAULT run shape through the drop-in (letf/config.py: NUM_SIMULATIONS = 200, all 19 strategies,
TIME_HORIZONS = [1, 2, 5, 10, 20, 30, 40, 50]): first pass (engine creation, scratch allocation) and a second pass.
usage: python scripts/default_run_latency.py [trials]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from letf_monte_carlo_b200 import parallel_monte_carlo_fixed, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ns = synthetic.config_namespace()
ns.NUM_SIMULATIONS = n
sids = list(ns.STRATEGIES)
model = synthetic.fallback_regime_model("default")
for label in ("first pass", "second pass"):
    t_all = time.perf_counter()
    for h in (1, 2, 5, 10, 20, 30, 40, 50):
        t0 = time.perf_counter()
        res = parallel_monte_carlo_fixed(sids, h, model, {}, cfg=ns, verbose=False)
        med = float(res["S4"].final_wealth.mean())
        torch.cuda.synchronize()
        print(f"{label}: horizon {h:2d} y  {1e3 * (time.perf_counter() - t0):8.1f} ms   mean S4 wealth {med:12.2f}   kernel {res.launch_info['kernel'] if hasattr(res, 'launch_info') else ''}")
    print(f"{label}: all horizons {1e3 * (time.perf_counter() - t_all):.1f} ms")
