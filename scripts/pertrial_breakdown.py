"""Where the time of the per-trial route (host result arrays) goes (tuning aid, GPU box).  usage: python scripts/pertrial_breakdown.py [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from letf_monte_carlo_b200 import mc_runner
from letf_monte_carlo_b200.outputs import HostOutputs
from letf_monte_carlo_b200.params import build_params

w = sys.argv[1] if len(sys.argv) > 1 else "c3"
ns, model, sids, params, keep = bench.build_workload(w)
n = bench.WORKLOADS[w]["trials"]; years = bench.WORKLOADS[w]["years"]
ns.NUM_SIMULATIONS = n
call = lambda: mc_runner.parallel_monte_carlo_fixed(sids, years, model, {}, cfg=ns, verbose=False)
for _ in range(3):
    r = call(); del r
def t(f, k=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3, r
ms, r = t(call); print(f"whole call (fresh result arrays)   {ms:8.2f} ms"); del r
params, keep = build_params(sids, years, model, ns)
eng = mc_runner._engine_for(params, keep, "f32", 0)
ms, out = t(lambda: eng.run_host(n, seed=1)); print(f"run_host, fresh HostOutputs        {ms:8.2f} ms   kernel {eng.launch_info()['last_kernel_ms']:.2f} ms")
ms, _ = t(lambda: eng.run_host(n, seed=1, out=out)); print(f"run_host, reused HostOutputs       {ms:8.2f} ms")
ms, _ = t(lambda: HostOutputs(len(sids), n)); print(f"HostOutputs allocation             {ms:8.2f} ms")
ms, dev = t(lambda: eng.run(n, seed=1)); print(f"run (device)                       {ms:8.2f} ms")
ms, _ = t(lambda: dev.to_host()); print(f"DeviceOutputs.to_host              {ms:8.2f} ms")
