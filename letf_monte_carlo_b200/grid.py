"""Horizon / leverage grids (BASELINE.json configs[4]): one Monte Carlo launch per horizon - the reference's own
semantics, `for horizon in horizons: parallel_monte_carlo_fixed(...)` (letf/__init__.py:150-162) - with the
leverages of a cell row simulated together as buy&hold series of one launch, per-cell terminal-wealth / CAGR /
max-drawdown histograms and moments reduced on the device (letf_reduce.cu) and allreduced across GPUs (NCCL).
Nothing per-trial leaves the device.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import numpy as np

from . import abi
from .distributed import exchange_summary, shard_range, world
from .params import build_params
from .summary import PERCENTILES, histogram_quantile


class HorizonGrid:
    """Engines and device buffers of a grid, reusable across repeated evaluations."""

    def __init__(self, horizons: Sequence[float], strategy_ids: Sequence[str], regime_model: Mapping, cfg, *,
                 precision="f32", device=None, lo=(-20.0, -1.0, 0.0), hi=(10.0, 2.0, 1.0)):
        from .engine import PathEngine, default_device
        self.horizons = [float(h) for h in horizons]
        self.strategy_ids = list(strategy_ids)
        self.device_index = default_device() if device is None else int(device)
        self.engines, self.summaries, self.params = [], [], []
        for h in self.horizons:
            p, keep = build_params(self.strategy_ids, h, regime_model, cfg)
            eng = PathEngine(p, keep, precision=precision, device=self.device_index)
            self.engines.append(eng)
            self.params.append(p)
            self.summaries.append(eng.new_summary(lo=lo, hi=hi,
                                                  bench_strategy=self.strategy_ids.index("S2") if "S2" in self.strategy_ids else -1))
        self.lo, self.hi = tuple(lo), tuple(hi)
        self._out = {}

    @property
    def launches(self):
        return sum(e.launches for e in self.engines)

    def path_days(self, n_trials: int) -> int:
        return sum(int(p.sim_days) * int(n_trials) for p in self.params)

    def run(self, n_trials: int, seed: int = 0, reduce_across_ranks: bool = True):
        """Simulate `n_trials` trials per horizon (sharded over the ranks) and fold them into the cell summaries."""
        rank, ws = world()
        for eng, sm, p in zip(self.engines, self.summaries, self.params):
            lo, hi = shard_range(n_trials, rank, ws, pair_aligned=bool(p.antithetic_pairs))
            n = hi - lo
            key = (id(eng), n)
            out = self._out.get(key)
            out = eng.run(n, first_trial=lo, seed=seed, out=out)
            self._out[key] = out
            eng.reset_summary(sm)
            eng.reduce(out, sm, n_trials=n)
            if reduce_across_ranks and ws > 1:
                exchange_summary(sm, engine=eng)
        return self

    def table(self, initial_capital: float = 10000.0) -> list[dict]:
        """One row per (horizon, strategy): wealth percentiles, median CAGR / max drawdown, counters."""
        rows = []
        for h, sm in zip(self.horizons, self.summaries):
            hist = sm.hist.cpu().numpy(); cnt = sm.counters.cpu().numpy(); mom = sm.moments.cpu().numpy()
            for s, sid in enumerate(self.strategy_ids):
                valid = int(cnt[s, 0])
                row = {"horizon": h, "id": sid, "valid": valid}
                for q in PERCENTILES:
                    row[f"p{q}"] = initial_capital * 10 ** histogram_quantile(hist[s, 0], self.lo[0], self.hi[0], q)
                row["cagr_median"] = histogram_quantile(hist[s, 1], self.lo[1], self.hi[1], 50)
                row["dd_median"] = histogram_quantile(hist[s, 2], self.lo[2], self.hi[2], 50)
                row["mean_log10_growth"] = float(mom[s, 0, 0]) / max(valid, 1)
                row["win_vs_S2"] = 100.0 * int(cnt[s, 1]) / max(valid, 1)
                rows.append(row)
        return rows

    def close(self):
        for e in self.engines:
            e.close()


assert abi.N_METRICS == 3 and np is not None
