"""Drop-in for `letf.mc_runner.parallel_monte_carlo_fixed` (reference: letf/mc_runner.py:24-210).

Same call shape and the same per-trial result mappings (`Final_Wealth`, `Max_DD`, `Num_Trades`,
`Trades_Per_Year`, `Trade_List`, `Metadata`), but the trial fan-out is one launch of the fused
CUDA path kernel instead of a loky process pool.  Results are lazy views over structure-of-arrays
buffers (at 1e7 trials a list of dicts is not an option, SURVEY.md 7 hard part 8).

Differences that are deliberate and documented:
  * `cfg` is snapshotted in the calling process at call time (the reference's workers re-import the
    config module and silently lose parent-side mutations, SURVEY.md 8b "Threading");
  * random numbers are counter-based Philox draws keyed by (cfg.GPU_SEED, sim_id), not numpy PCG64
    streams: distributions agree (KS-tested), individual paths do not;
  * `Regime_Path` (a T-element list per result, never read downstream) is replaced by the regime
    day counts in `Metadata['regime_counts']`.
"""
from __future__ import annotations

from collections.abc import Mapping, Sequence

import numpy as np

from . import abi
from . import config as _default_cfg
from .distributed import shard_range, world
from .params import build_params
from .trade import TAXABLE_IDS


class TrialResult(Mapping):
    """One trial of one strategy; behaves like the reference's result dict (engine.py:592-603)."""
    _KEYS = ("Final_Wealth", "Num_Trades", "Trades_Per_Year", "Max_DD", "Metadata", "Trade_List")

    def __init__(self, parent: "StrategyResults", i: int):
        self._p, self._i = parent, i

    def __getitem__(self, k):
        p, i = self._p, self._i
        if k == "Final_Wealth":
            return float(p.final_wealth[i])
        if k == "Max_DD":
            return float(p.max_dd[i])
        if k == "Num_Trades":
            return int(p.num_trades[i])
        if k == "Trades_Per_Year":
            return float(p.num_trades[i]) / p.run.years
        if k == "Metadata":
            return p.run.metadata(i)
        if k == "Trade_List" and p.sid in p.run.taxable_ids:
            return p.trade_list(i)
        raise KeyError(k)

    def __iter__(self):
        return (k for k in self._KEYS if k != "Trade_List" or self._p.sid in self._p.run.taxable_ids)

    def __len__(self):
        return sum(1 for _ in self)


class StrategyResults(Sequence):
    """`all_results[sid]`: sequence of per-trial mappings + array accessors."""

    def __init__(self, run: "MonteCarloRun", sid: str, s: int):
        self.run, self.sid, self.s = run, sid, s
        self.final_wealth = run.out.final_wealth[s]
        self.max_dd = run.out.max_dd[s]
        self.num_trades = run.out.num_trades[s]

    def __len__(self):
        return self.run.n_local

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return TrialResult(self, i)

    def trade_list(self, i: int):
        return self.run.engine.trade_list(self.run.first_trial + i, self.s, seed=self.run.seed)

    def trade_lists(self, idx):
        """several trials in one launch (what the sampled-trial tax path asks for)"""
        return self.run.engine.trade_lists([self.run.first_trial + int(i) for i in idx], self.s, seed=self.run.seed)


class MonteCarloRun(dict):
    """The returned mapping {sid: StrategyResults}; also keeps the launch metadata."""

    def __init__(self, engine, out, strategy_ids, years, seed, first_trial, n_local, n_total, taxable_ids):
        super().__init__()
        self.engine, self.out, self.years, self.seed = engine, out, float(years), seed
        self.first_trial, self.n_local, self.n_total = first_trial, n_local, n_total
        self.taxable_ids = set(taxable_ids)
        for s, sid in enumerate(strategy_ids):
            self[sid] = StrategyResults(self, sid, s)

    def taxable_summary(self, tax_fn=None, tax_config=None, cfg=None, **kw):
        """Rows of the reference's taxable-brokerage table (reporting.py:399-560): <= 50 wealth-ranked trials per
        strategy are replayed on the device for their Trade_List and taxed by `tax_fn` (tax_sampling.py)."""
        from .tax_sampling import taxable_summary
        cfg = _default_cfg if cfg is None else cfg
        return taxable_summary(self, self.years, tax_fn=tax_fn, tax_config=tax_config,
                               initial_capital=float(cfg.INITIAL_CAPITAL), names=cfg.STRATEGIES,
                               taxable_ids=[s for s in TAXABLE_IDS if s in self], **kw)

    def metadata(self, i: int) -> dict:
        o = self.out
        fl = int(o.flags[i])
        return {"layer_checks": {"is_valid": not (fl & abi.FLAG_REPAIRED),
                                 "repaired_after_validation": bool(fl & abi.FLAG_REPAIRED)},
                "regime_counts": {0: int(o.regime_days[0, i]), 1: int(o.regime_days[1, i])},
                "stress_summary": {"jump_days": int(o.jump_days[i])},
                "start_offset": int(o.start_offset[i]),
                "sim_id": int((self.first_trial + i) & ~1) if self.engine.params.antithetic_pairs else int(self.first_trial + i),
                "engine": "letf_monte_carlo_b200"}


_ENGINES: dict = {}      # (parameter bytes, precision, device) -> PathEngine; the handle owns device scratch worth reusing
_ENGINE_CACHE_MAX = 8


def _engine_for(params, keep, precision, device):
    """Engines are pure functions of their parameter block, so repeated calls (one per horizon in
    letf.run(), letf/__init__.py:150-162) reuse the handle instead of re-creating device state."""
    import ctypes as C
    from .engine import PathEngine, default_device
    dev = default_device() if device is None else int(device)
    probe = type(params)()
    C.memmove(C.byref(probe), C.byref(params), C.sizeof(params))
    durs = []
    for r in range(2):
        nd = int(params.n_duration[r])
        durs.append(bytes(C.string_at(params.duration_samples[r], 4 * nd)) if nd > 0 else b"")
        probe.duration_samples[r] = None                   # host pointers are not part of the identity
    if params.market_mode == abi.MARKET_BOOTSTRAP:            # same for the bootstrap's host tables
        import hashlib
        h = hashlib.sha256()
        b = params.boot
        for ptr in (b.hist_spy, b.hist_qqq, b.hist_tlt):
            h.update(C.string_at(ptr, 8 * b.hist_len))
        for pools, probe_pools in ((b.pool_a, probe.boot.pool_a), (b.pool_b, probe.boot.pool_b)):
            for r in range(2):
                h.update(C.string_at(pools[r].start, 4 * pools[r].n_blocks))
                h.update(C.string_at(pools[r].block_ret, 8 * pools[r].n_blocks))
                probe_pools[r].start = None
                probe_pools[r].block_ret = None
        probe.boot.hist_spy = probe.boot.hist_qqq = probe.boot.hist_tlt = None
        durs.append(h.digest())
    key = (bytes(probe), tuple(durs), str(precision), dev)
    eng = _ENGINES.get(key)
    if eng is None or eng._h is None:
        while len(_ENGINES) >= _ENGINE_CACHE_MAX:
            _ENGINES.pop(next(iter(_ENGINES))).close()
        eng = _ENGINES[key] = PathEngine(params, keep, precision=precision, device=dev)
    return eng


def parallel_monte_carlo_fixed(strategy_ids, time_horizon, regime_model, correlation_matrices,
                               historical_df=None, *, cfg=None, precision=None, seed=None, device=None,
                               verbose=True):
    """Reference signature: letf/mc_runner.py:24-25.  `correlation_matrices` is accepted and unused,
    exactly as in the reference (strategy.py:55 is its only reader and ignores it)."""
    from .engine import PathEngine
    cfg = _default_cfg if cfg is None else cfg
    if historical_df is not None and getattr(cfg, "USE_BLOCK_BOOTSTRAP", False):
        # mc_runner.py:79-84: the sampler rides in the (shallow-copied) regime_model; the generator dispatch
        # (engine.py:45-52) only reaches it when the engine mode is not institutional_v1 with a joint model
        from .bootstrap import create_bootstrap_sampler
        regime_model = dict(regime_model)
        regime_model["bootstrap_sampler"] = create_bootstrap_sampler(historical_df, cfg=cfg)
    if historical_df is not None and getattr(cfg, "USE_RANDOM_START", False) \
            and getattr(cfg, "RANDOM_START_METHOD", "") == "historical_anchor":
        regime_model = dict(regime_model)                      # mc_runner.py:120-124
        regime_model["historical_df_for_anchors"] = historical_df
    strategy_ids = list(strategy_ids)
    n_total = int(cfg.NUM_SIMULATIONS)
    params, keep = build_params(strategy_ids, time_horizon, regime_model, cfg)
    rank, ws = world()
    lo, hi = shard_range(n_total, rank, ws, pair_aligned=bool(params.antithetic_pairs))
    precision = precision or getattr(cfg, "GPU_PRECISION", "f32")
    seed = int(getattr(cfg, "GPU_SEED", 0) if seed is None else seed)
    if verbose and rank == 0:
        print(f"\n{'=' * 80}\nMONTE CARLO: {n_total:,} sims x {time_horizon}Y  "
              f"[B200 path engine, {precision}, {ws} GPU(s)]\n{'=' * 80}")
    eng = _engine_for(params, keep, precision, device if device is not None else getattr(cfg, "GPU_DEVICE", None))
    # raises NonFiniteTrials like the reference's ValueError (the C-ABI stages through its own pinned buffer)
    out = eng.run_host(hi - lo, first_trial=lo, seed=seed)
    return MonteCarloRun(eng, out, strategy_ids, time_horizon, seed, lo, hi - lo, n_total, TAXABLE_IDS)
