"""Where the time of the sampled-trial tax table goes (tuning aid, GPU box): Monte Carlo, device Trade_List replay, host
conversion, host tax engine.  usage: python scripts/tax_path_breakdown.py [trials] [years]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from letf_monte_carlo_b200 import parallel_monte_carlo_fixed, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
years = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ns = synthetic.config_namespace()
ns.NUM_SIMULATIONS = n
sids = list(ns.STRATEGIES)
model = synthetic.fallback_regime_model("default")
res = parallel_monte_carlo_fixed(sids, years, model, {}, cfg=ns, verbose=False)
t0 = time.perf_counter(); res = parallel_monte_carlo_fixed(sids, years, model, {}, cfg=ns, verbose=False); t1 = time.perf_counter()
print(f"Monte Carlo ({n} trials x {years} y x {len(sids)} strategies)  {1e3 * (t1 - t0):9.1f} ms")
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); rows = res.taxable_summary(); t1 = time.perf_counter()
pr.disable()
print(f"taxable_summary ({len(rows)} strategies)                  {1e3 * (t1 - t0):9.1f} ms")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
