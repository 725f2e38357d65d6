"""Where the end-to-end time of parallel_monte_carlo_fixed goes (tuning aid, GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from letf_monte_carlo_b200 import synthetic, mc_runner
from letf_monte_carlo_b200.params import build_params
from letf_monte_carlo_b200.outputs import HostOutputs

ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False, USE_ANTITHETIC_VARIATES=False)
ns.NUM_SIMULATIONS = 100000
sids = ["S1", "B_UPRO", "S3"]; model = synthetic.normal_regime_model("c2")
for _ in range(3):
    mc_runner.parallel_monte_carlo_fixed(sids, 30, model, {}, cfg=ns, verbose=False)
def t(f, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
ms, _ = t(lambda: mc_runner.parallel_monte_carlo_fixed(sids, 30, model, {}, cfg=ns, verbose=False)); print(f"whole call        {ms:8.3f} ms")
ms, pk = t(lambda: build_params(sids, 30, model, ns)); print(f"build_params      {ms:8.3f} ms")
params, keep = pk
ms, eng = t(lambda: mc_runner._engine_for(params, keep, "f32", 0)); print(f"engine lookup     {ms:8.3f} ms")
ms, out = t(lambda: HostOutputs(3, 100000, journal=False, pinned=True)); print(f"pinned outputs    {ms:8.3f} ms")
ms, _ = t(lambda: eng.run_host(100000, seed=1, out=out)); print(f"run_host (reuse)  {ms:8.3f} ms")
ms, d = t(lambda: eng.run(100000, seed=1)); print(f"run (device)      {ms:8.3f} ms")
ms, _ = t(lambda: d.to_host()); print(f"to_host pageable  {ms:8.3f} ms")
