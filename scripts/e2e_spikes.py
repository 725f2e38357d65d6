"""Call-to-call jitter of the end-to-end route (tuning aid, GPU box): 40 calls with and without the cyclic collector,
and the same number of bare engine.run + synchronize calls.  usage: python scripts/e2e_spikes.py [workload] [trials]"""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from letf_monte_carlo_b200 import mc_runner
from letf_monte_carlo_b200.params import build_params

w = sys.argv[1] if len(sys.argv) > 1 else "c4"
ns, model, sids, params, keep = bench.build_workload(w)
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[w]["trials"]; years = bench.WORKLOADS[w]["years"]
ns.NUM_SIMULATIONS = n
call = lambda: mc_runner.parallel_monte_carlo_fixed(sids, years, model, {}, cfg=ns, verbose=False, summary_only=True).table()
for _ in range(3):
    call()
def series(f, k=40):
    out = []
    for _ in range(k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3)
    return np.array(out)
def show(tag, x):
    print(f"{tag:28s} median {np.median(x):8.2f}  max {x.max():8.2f}  mean {x.mean():8.2f}  >1.1x median: {int((x > 1.1 * np.median(x)).sum())}/{len(x)}   gc counts {gc.get_count()}")
show("call, gc on", series(call))
gc.disable()
show("call, gc off", series(call))
gc.enable()
params, keep = build_params(sids, years, model, ns)
eng = mc_runner._engine_for(params, keep, "f32", 0)
dev = eng.run(n, seed=1)
show("engine.run (reuse bufs)", series(lambda: eng.run(n, seed=1, out=dev)))
km = []
def krun():
    eng.run(n, seed=1, out=dev); torch.cuda.synchronize(); km.append(eng.launch_info()["last_kernel_ms"])
series(krun)
show("kernel ms (events)", np.array(km))
host, wait = [], []
for _ in range(40):
    torch.cuda.synchronize(); t0 = time.perf_counter(); eng.run(n, seed=1, out=dev); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    host.append((t1 - t0) * 1e3); wait.append((t2 - t1) * 1e3)
show("engine.run returns (host)", np.array(host))
show("synchronize after it", np.array(wait))
print("host:", " ".join(f"{x:.1f}" for x in host))
