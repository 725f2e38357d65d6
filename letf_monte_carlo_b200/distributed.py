"""Trial sharding and the one exchange step of the path.

Trials are independent pure functions of (seed, trial id), so ranks simulate disjoint
contiguous trial ranges with no communication; the only collective is the reduction of the
summary buffers (histogram bins, moments, counters: SUM; extrema: MIN / MAX) at the end.
`torch.distributed` is the plumbing (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

import os


def world():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_trials: int, rank: int, world_size: int, pair_aligned: bool = True):
    """Contiguous trial range [lo, hi) of `rank`.  With antithetic pairing the boundaries are kept
    even so that twins (2k, 2k+1), which share their sim_id, stay on one GPU."""
    unit = 2 if pair_aligned else 1
    units = (n_trials + unit - 1) // unit
    base, extra = divmod(units, world_size)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(lo_u * unit, n_trials), min(hi_u * unit, n_trials)


def allreduce_summary(hist, moments, extrema, counters, group=None):
    """In-place reduction of the letf_reduce buffers across ranks.
    extrema[..., 0] holds minima, extrema[..., 1] maxima."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    import torch
    # bins and counters are both int64 SUM: one collective on a packed buffer
    packed_i = torch.cat([hist.reshape(-1), counters.reshape(-1)])
    dist.all_reduce(packed_i, op=dist.ReduceOp.SUM, group=group)
    hist.copy_(packed_i[:hist.numel()].view_as(hist))
    counters.copy_(packed_i[hist.numel():].view_as(counters))
    dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    # one MAX collective for both extrema: negate the minima
    packed = extrema.clone()
    packed[..., 0].neg_()
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
    packed[..., 0].neg_()
    extrema.copy_(packed)


def exchange_summary(summary, engine=None, group=None):
    """The exchange step on a packed `DeviceSummary`: ONE all-gather of the packed buffer (NCCL over NVLink) and one
    combine kernel (`letf_summary_combine`: integer sums for bins / counters, fp64 sums for the moments, min / max for the
    extrema) - instead of three all-reduces plus packing kernels.  CPU tensors (the gloo tests of the host logic) are
    combined with torch ops."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return summary
    ws = dist.get_world_size(group)
    buf = summary.buf
    gathered = torch.empty((ws, buf.numel()), dtype=buf.dtype, device=buf.device)
    if buf.is_cuda:
        dist.all_gather_into_tensor(gathered, buf, group=group)
        if engine is None:
            raise RuntimeError("exchange_summary on device tensors needs the PathEngine (letf_summary_combine)")
        engine.combine_summaries(gathered, ws, summary)
        return summary
    dist.all_gather(list(gathered.unbind(0)), buf, group=group)
    nh, nc = summary.hist.numel(), summary.counters.numel()
    nm = summary.moments.numel()
    buf[:nh + nc] = gathered[:, :nh + nc].sum(0)
    f = gathered[:, nh + nc:].view(torch.float64)
    buf[nh + nc:nh + nc + nm] = f[:, :nm].sum(0).view(torch.int64)
    ex = f[:, nm:].view(ws, -1, 2)
    out = torch.stack([ex[:, :, 0].amin(0), ex[:, :, 1].amax(0)], dim=1).reshape(-1)
    buf[nh + nc + nm:] = out.view(torch.int64)
    return summary
