"""Where the end-to-end time of parallel_monte_carlo_fixed(..., summary_only=True).table() goes (tuning aid, GPU box).
usage: python scripts/e2e_breakdown.py [workload] [trials]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from letf_monte_carlo_b200 import mc_runner
from letf_monte_carlo_b200.params import build_params

w = sys.argv[1] if len(sys.argv) > 1 else "c3"
ns, model, sids, params, keep = bench.build_workload(w)
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[w]["trials"]; years = bench.WORKLOADS[w]["years"]
ns.NUM_SIMULATIONS = n
call = lambda: mc_runner.parallel_monte_carlo_fixed(sids, years, model, {}, cfg=ns, verbose=False, summary_only=True).table()
for _ in range(3):
    call()
each = []
for _ in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter(); call(); torch.cuda.synchronize(); each.append((time.perf_counter() - t0) * 1e3)
print("calls (ms):", " ".join(f"{x:.1f}" for x in each))
def t(f, k=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3, r
ms, _ = t(call); print(f"whole call              {ms:8.3f} ms")
ms, pk = t(lambda: build_params(sids, years, model, ns)); print(f"build_params            {ms:8.3f} ms")
params, keep = pk
ms, eng = t(lambda: mc_runner._engine_for(params, keep, "f32", 0)); print(f"engine lookup           {ms:8.3f} ms")
ms, dev = t(lambda: eng.run(n, seed=1)); print(f"run (device, new bufs)  {ms:8.3f} ms   kernel {eng.launch_info()['last_kernel_ms']:.3f} ms")
ms, dev = t(lambda: eng.run(n, seed=1, out=dev)); print(f"run (device, reuse)     {ms:8.3f} ms")
ms, sm = t(lambda: eng.new_summary(bench_strategy=0)); print(f"new_summary             {ms:8.3f} ms")
ms, _ = t(lambda: eng.reduce(dev, sm)); print(f"reduce                  {ms:8.3f} ms")
ms, b = t(lambda: sm.buf.cpu()); print(f"summary D2H             {ms:8.3f} ms")
run = mc_runner.parallel_monte_carlo_fixed(sids, years, model, {}, cfg=ns, verbose=False, summary_only=True)
run._fetch()
ms, _ = t(lambda: run.table()); print(f"table() from host copy  {ms:8.3f} ms")
