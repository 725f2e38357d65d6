"""Multi-rank check of the exact cross-rank percentiles on NCCL (run under torchrun on >= 2 GPUs):
MonteCarloSummary.table(exact=True) on sharded device arrays == np.percentile of the gathered arrays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from letf_monte_carlo_b200 import parallel_monte_carlo_fixed, synthetic
from letf_monte_carlo_b200.summary import summary_table

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ns = synthetic.config_namespace(NUM_SIMULATIONS=30001 * 2, GPU_SEED=5)
sids = ["S1", "S2", "S4", "S19"]
model = synthetic.fallback_regime_model("default")
sm = parallel_monte_carlo_fixed(sids, 2, model, {}, cfg=ns, verbose=False, summary_only=True, device=local)
rows = {r["id"]: r for r in sm.table(exact=True)}
# reference: gather every rank's wealth and use numpy
ok = True
for s, sid in enumerate(sids):
    mine = sm.device_out.final_wealth[s].contiguous()
    sizes = [torch.zeros(1, dtype=torch.int64, device=mine.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device))
    m = int(max(int(x) for x in sizes))
    pad = torch.zeros(m, dtype=mine.dtype, device=mine.device); pad[:mine.numel()] = mine
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    allw = torch.cat([p[:int(n)] for p, n in zip(parts, sizes)]).cpu().numpy()
    for q in (10, 25, 40, 50, 60, 75, 90):
        ok &= rows[sid][f"p{q}"] == float(np.percentile(allw, q))
    ok &= rows[sid]["valid"] == len(allw) == ns.NUM_SIMULATIONS
# the array-backed table on a sharded per-trial run agrees too
full = parallel_monte_carlo_fixed(sids, 2, model, {}, cfg=ns, verbose=False, device=local)
t2 = {r["id"]: r for r in summary_table(full, 2)}
for sid in sids:
    ok &= t2[sid]["median"] == rows[sid]["p50"] and t2[sid]["p90"] == rows[sid]["p90"]
flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local))
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("exact percentiles across", ws, "ranks:", "OK" if int(flag) else "MISMATCH")
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
