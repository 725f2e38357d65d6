"""Post-Monte-Carlo statistics in the reference's definitions (letf/reporting.py:353-388, 320-325, 541-555),
computed from the engine's result arrays instead of per-trial Python dicts.

`summary_table` is the array-backed equivalent of the rows `create_summary_statistics` assembles before it
prints: exact `np.percentile` (linear interpolation) at P10/25/40/60/75/90 and the median of Final_Wealth,
`CAGR = (W / INITIAL_CAPITAL) ** (1 / years) - 1`, median Max_DD, mean Trades_Per_Year and the win rate
against S2 **by trial index**.  `histogram_quantile` reads the same percentiles off the allreduced
4096-bin histograms of `letf_reduce` (the multi-GPU summary; bin-width accuracy).
"""
from __future__ import annotations

from typing import Mapping

import numpy as np

from . import abi

PERCENTILES = (10, 25, 40, 50, 60, 75, 90)


def _dist_group(group=None):
    """(torch.distributed module, world size) when a multi-rank run is active, else (None, 1)"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist, dist.get_world_size(group)
    except Exception:
        pass
    return None, 1


def _comm_tensor(t, group=None):
    """`t` on the device the process group's backend communicates on (NCCL: this rank's GPU; gloo: as is)"""
    import torch
    import torch.distributed as dist
    if not t.is_cuda and "nccl" in str(dist.get_backend(group)):
        return t.to(torch.device("cuda", torch.cuda.current_device()))
    return t


def _np_lerp(a: float, b: float, t: float) -> float:
    """numpy's 'linear' percentile interpolation (numpy/lib/_function_base_impl.py _lerp), bit for bit"""
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def global_order_statistics(values, ks, group=None, bins: int = 4096, cap: int = 1 << 16):
    """Exact order statistics of the UNION of every rank's `values` (1-D torch tensor, CPU or CUDA) without gathering them:
    histogram-guided selection.  Per requested global rank k: all ranks histogram their values over the
    current bracket (one SUM all-reduce of `bins` counters), the bin that holds rank k becomes the new bracket, and once it
    holds <= cap values globally those candidates are gathered (a second, tiny exchange), sorted and indexed.  Returns
    floats, identical on every rank."""
    import torch
    dist, ws = _dist_group(group)
    v = values.reshape(-1).to(torch.float64)
    if dist is None:
        sv = torch.sort(v).values
        return [float(sv[int(k)]) for k in ks]
    v = _comm_tensor(v, group)

    def allred(t, op):
        dist.all_reduce(t, op=op, group=group)
        return t

    def gather_sorted(c):
        cnt = torch.tensor([c.numel()], dtype=torch.int64, device=v.device)
        cnts = [torch.zeros_like(cnt) for _ in range(ws)]
        dist.all_gather(cnts, cnt, group=group)
        m = int(max(int(x) for x in cnts))
        pad = torch.full((max(m, 1),), float("inf"), dtype=torch.float64, device=v.device)
        pad[:c.numel()] = c
        parts = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(parts, pad, group=group)
        return torch.sort(torch.cat([p[:int(n)] for p, n in zip(parts, cnts)])).values

    out = []
    for k in ks:
        sub, below = v, 0
        for _ in range(16):
            fin = sub if sub.numel() else torch.zeros(1, dtype=torch.float64, device=v.device)
            a = float(allred(fin.min().clone() if sub.numel() else torch.tensor(float("inf"), dtype=torch.float64, device=v.device),
                             dist.ReduceOp.MIN))
            b = float(allred(fin.max().clone() if sub.numel() else torch.tensor(float("-inf"), dtype=torch.float64, device=v.device),
                             dist.ReduceOp.MAX))
            total = int(allred(torch.tensor([sub.numel()], dtype=torch.int64, device=v.device), dist.ReduceOp.SUM))
            if not (b > a) or total <= cap:
                if b > a or total <= cap:
                    out.append(float(gather_sorted(sub)[int(k) - below]))
                else:
                    out.append(a)
                break
            idx = torch.clamp(((sub - a) * (bins / (b - a))).floor().to(torch.int64), 0, bins - 1)
            h = allred(torch.bincount(idx, minlength=bins), dist.ReduceOp.SUM)
            cum = torch.cumsum(h, 0)
            t = int(k) - below
            bsel = int(torch.searchsorted(cum, torch.tensor([t], dtype=cum.dtype, device=cum.device), right=True))
            below += int(cum[bsel] - h[bsel])
            sub = sub[idx == bsel]
        else:
            raise RuntimeError("global_order_statistics did not converge")
    return out


def global_percentiles(values, qs, group=None) -> list[float]:
    """`np.percentile(union of all ranks' values, qs)` (linear interpolation, reporting.py:361-366), exact: the two order
    statistics around each percentile come from `global_order_statistics`, the interpolation is numpy's."""
    import torch
    dist, ws = _dist_group(group)
    v = values if isinstance(values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(values))
    if dist is None:
        return [float(x) for x in np.percentile(v.detach().cpu().numpy().astype(np.float64), list(qs))]
    v = _comm_tensor(v, group)
    n = torch.tensor([v.numel()], dtype=torch.int64, device=v.device)
    dist.all_reduce(n, group=group)
    n = int(n)
    pos = [q / 100.0 * (n - 1) for q in qs]
    need = sorted({int(np.floor(p)) for p in pos} | {int(np.ceil(p)) for p in pos})
    stat = dict(zip(need, global_order_statistics(v, need, group=group)))
    return [_np_lerp(stat[int(np.floor(p))], stat[int(np.ceil(p))], p - np.floor(p)) for p in pos]


def _arrays(results):
    """(final_wealth, max_dd, num_trades) from a StrategyResults view or a list of reference-style dicts."""
    if hasattr(results, "final_wealth"):
        return np.asarray(results.final_wealth), np.asarray(results.max_dd), np.asarray(results.num_trades), None
    fw = np.array([r["Final_Wealth"] for r in results], dtype=float)
    dd = np.array([r.get("Max_DD", 0) for r in results], dtype=float)
    tpy = np.array([r.get("Trades_Per_Year", 0) for r in results], dtype=float)
    return fw, dd, None, tpy


def summary_table(mc_results: Mapping, time_horizon: float, initial_capital: float = 10000.0,
                  benchmark_id: str = "S2", names: Mapping | None = None) -> list[dict]:
    """One row per strategy, sorted by median wealth (descending) like the reference's table."""
    rows = []
    spy = None
    if benchmark_id in mc_results and len(mc_results[benchmark_id]):
        spy = _arrays(mc_results[benchmark_id])[0]
    dist, ws = _dist_group()
    sharded = dist is not None and getattr(mc_results, "n_total", None) not in (None, getattr(mc_results, "n_local", None))
    for sid, results in mc_results.items():
        if not len(results) and not sharded:
            continue
        fw, dd, nt, tpy = _arrays(results)
        if sharded:
            # this rank holds a shard: exact percentiles of the union by histogram-guided selection, means by all-reduce
            import torch
            qs = [q for q in PERCENTILES if q != 50] + [50]
            vals = global_percentiles(torch.from_numpy(np.ascontiguousarray(fw, dtype=np.float64)), qs)
            pct = {f"p{q}": v for q, v in zip(qs[:-1], vals[:-1])}
            median = vals[-1]
            tsum = float(np.sum(nt)) / time_horizon if nt is not None else float(np.sum(tpy))
            acc = _comm_tensor(torch.tensor([float(np.sum(fw > spy)) if spy is not None and len(spy) == len(fw) else 0.0, tsum,
                                             float(len(fw))], dtype=torch.float64))
            dist.all_reduce(acc)
            win, trades = float(acc[0] / acc[2] * 100), float(acc[1] / acc[2])
            ddm = global_percentiles(torch.from_numpy(np.ascontiguousarray(dd, dtype=np.float64)), [50])[0]
            rows.append({"id": sid, "name": (names or {}).get(sid, {}).get("name", sid) if names else sid,
                         **pct, "median": median, "cagr": (median / initial_capital) ** (1 / time_horizon) - 1,
                         "win": win, "dd": ddm, "trades": trades})
            continue
        pct = {f"p{q}": float(np.percentile(fw, q)) for q in PERCENTILES if q != 50}
        median = float(np.median(fw))
        win = float(np.mean(fw > spy) * 100) if spy is not None and len(spy) == len(fw) else 0.0
        trades = float(np.mean(nt) / time_horizon) if nt is not None else float(np.mean(tpy))
        rows.append({"id": sid, "name": (names or {}).get(sid, {}).get("name", sid) if names else sid,
                     **pct, "median": median, "cagr": (median / initial_capital) ** (1 / time_horizon) - 1,
                     "win": win, "dd": float(np.median(dd)), "trades": trades})
    rows.sort(key=lambda r: r["median"], reverse=True)
    return rows


def cagr(wealth, time_horizon: float, initial_capital: float = 10000.0):
    return (np.asarray(wealth, dtype=float) / initial_capital) ** (1 / time_horizon) - 1


def histogram_quantile(hist: np.ndarray, lo: float, hi: float, q: float) -> float:
    """q-th percentile (0..100) of a letf_reduce histogram (bins equally spaced on [lo, hi]); linear inside the bin."""
    hist = np.asarray(hist, dtype=np.float64)
    n = hist.sum()
    if n <= 0:
        return float("nan")
    target = q / 100.0 * n
    c = np.cumsum(hist)
    b = int(np.searchsorted(c, target, side="left"))
    b = min(b, len(hist) - 1)
    prev = c[b - 1] if b > 0 else 0.0
    frac = (target - prev) / hist[b] if hist[b] > 0 else 0.0
    width = (hi - lo) / len(hist)
    return lo + (b + frac) * width


def summary_from_histograms(hist, lo, hi, counters, initial_capital: float = 10000.0) -> list[dict]:
    """Per strategy: wealth percentiles from the log10(W/W0) histogram, median max drawdown, counters.
    hist: [S][3][bins] (metric 0 log10(W/W0), 1 CAGR, 2 max drawdown), counters: [S][4]."""
    out = []
    for s in range(hist.shape[0]):
        row = {f"p{q}": initial_capital * 10 ** histogram_quantile(hist[s, 0], lo[0], hi[0], q) for q in PERCENTILES}
        row["cagr_median"] = histogram_quantile(hist[s, 1], lo[1], hi[1], 50)
        row["dd_median"] = histogram_quantile(hist[s, 2], lo[2], hi[2], 50)
        valid = int(counters[s, 0])
        row.update(valid=valid, win=100.0 * int(counters[s, 1]) / max(valid, 1), total_trades=int(counters[s, 2]),
                   flagged=int(counters[s, 3]))
        out.append(row)
    return out


assert abi.N_METRICS == 3
