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
    """One trial of one strategy; behaves like the reference's result dict (engine.py:592-603).

    A strategy whose equity ended non-finite is reported the way the reference reports a strategy that raised inside
    `run_strategy_fixed` (engine.py:606-617): a zero-wealth record carrying an 'Error' key, so that downstream
    `r['Final_Wealth']` / `r.get('Max_DD', 0)` keep working and the failure stays visible.  (The reference's 300 s
    per-trial timeout and its `None`-skip, mc_runner.py:169,197-208, have no counterpart: a launch either completes
    or raises.)"""
    _KEYS = ("Final_Wealth", "Num_Trades", "Trades_Per_Year", "Max_DD", "Metadata", "Trade_List")

    def __init__(self, parent: "StrategyResults", i: int):
        self._p, self._i = parent, i

    def _failed(self):
        return not np.isfinite(self._p.final_wealth[self._i])

    def __getitem__(self, k):
        p, i = self._p, self._i
        if self._failed():
            if k == "Error":
                return "non-finite strategy equity"
            if k in ("Final_Wealth", "Max_DD", "Trades_Per_Year"):
                return 0.0
            if k == "Num_Trades":
                return 0
            if k == "Trade_List" and p.sid in p.run.taxable_ids:
                return []
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
        keys = [k for k in self._KEYS if k != "Trade_List" or self._p.sid in self._p.run.taxable_ids]
        return iter(keys + (["Error"] if self._failed() else []))

    def __len__(self):
        return sum(1 for _ in self)


class StrategyResults(Sequence):
    """`all_results[sid]`: sequence of per-trial mappings + array accessors."""

    def __init__(self, run: "MonteCarloRun", sid: str, s: int):
        self.run, self.sid, self.s = run, sid, s

    # host views; the device -> host copy of the per-trial arrays happens on first access (MonteCarloRun.out)
    @property
    def final_wealth(self):
        return self.run.out.final_wealth[self.s]

    @property
    def max_dd(self):
        return self.run.out.max_dd[self.s]

    @property
    def num_trades(self):
        return self.run.out.num_trades[self.s]

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
    """The returned mapping {sid: StrategyResults}; also keeps the launch metadata.

    `out` is either the host arrays `letf_run_host` filled, or - `device_out` given - fetched from the device on first
    access: a caller that only wants `summary()` never pays the per-trial device -> host copy."""

    def __init__(self, engine, out, strategy_ids, years, seed, first_trial, n_local, n_total, taxable_ids, device_out=None):
        super().__init__()
        self.engine, self._out, self.device_out = engine, out, device_out
        self.years, self.seed = float(years), seed
        self.first_trial, self.n_local, self.n_total = first_trial, n_local, n_total
        self.taxable_ids = set(taxable_ids)
        self.strategy_ids = list(strategy_ids)
        self._summary = None
        # the mapping stores slot indices and builds the (cheap) StrategyResults views on access: a stored view would point
        # back at the run - a reference cycle, so a dropped run (and its device buffers) would wait for the cyclic collector
        # and every call would have to cudaMalloc fresh result buffers
        for s, sid in enumerate(strategy_ids):
            dict.__setitem__(self, sid, s)

    def __getitem__(self, sid):
        return StrategyResults(self, sid, dict.__getitem__(self, sid))

    def get(self, sid, default=None):
        return self[sid] if sid in self else default

    def values(self):
        return [self[k] for k in self]

    def items(self):
        return [(k, self[k]) for k in self]

    @property
    def out(self):
        if self._out is None:
            self._out = self.device_out.to_host()
            self.engine.raise_if_nonfinite(self._out.flags)
        return self._out

    def summary(self, benchmark_id: str = "S2"):
        """`MonteCarloSummary` of this run: reduced on the device (letf_reduce) and exchanged across the ranks."""
        if self._summary is None:
            if self.device_out is None:
                raise RuntimeError("summary() needs the device-resident route: parallel_monte_carlo_fixed(..., summary_only=True) "
                                   "or device_results=True")
            self._summary = MonteCarloSummary(self, benchmark_id)
        return self._summary

    def trade_lists_multi(self, wanted):
        """{sid: [trial indices]} -> {sid: [Trade_List ...]} in ONE device launch (engine.trade_lists_multi)"""
        sids = [sid for sid in wanted if len(wanted[sid])]
        trials = [self.first_trial + int(i) for sid in sids for i in wanted[sid]]
        strategies = [self[sid].s for sid in sids for _ in wanted[sid]]
        flat = self.engine.trade_lists_multi(trials, strategies, seed=self.seed)
        out, k = {sid: [] for sid in wanted}, 0
        for sid in sids:
            out[sid] = flat[k:k + len(wanted[sid])]
            k += len(wanted[sid])
        return out

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


class MonteCarloSummary:
    """Summary statistics of a run, reduced on the device: per (strategy, metric) 4096-bin histograms, moments, extrema and
    counters (`letf_reduce`), exchanged across the ranks with one all-gather + one combine kernel.  Only these buffers
    (~0.1 MB per strategy) ever leave the device; `table()` reads the reference's report columns off them
    (reporting.py:353-388): histogram percentiles (bin-width accuracy) or, with exact=True, np.percentile-exact values of
    the union of all ranks' trials by histogram-guided selection on the device-resident per-trial arrays."""

    def __init__(self, run: MonteCarloRun, benchmark_id: str = "S2"):
        from .distributed import exchange_summary
        eng = run.engine
        # (no reference back to the run: see MonteCarloRun.__init__)
        self.device_out, self.strategy_ids, self.years, self.n_total = run.device_out, list(run.strategy_ids), run.years, run.n_total
        bench = self.strategy_ids.index(benchmark_id) if benchmark_id in self.strategy_ids else -1
        self.sm = eng.new_summary(bench_strategy=bench)
        eng.reduce(self.device_out, self.sm, n_trials=run.n_local)
        exchange_summary(self.sm, engine=eng)
        self._host = None

    def _fetch(self):
        if self._host is None:
            import torch
            b = self.sm.buf.cpu()                                       # the one device -> host copy of this route
            nh, nc = self.sm.hist.numel(), self.sm.counters.numel()
            nm = self.sm.moments.numel()
            S = max(self.sm.n_strategies, 1)
            f = b[nh + nc:].view(torch.float64).numpy()
            self._host = {"hist": b[:nh].numpy().reshape(S, abi.N_METRICS, abi.HIST_BINS), "counters": b[nh:nh + nc].numpy().reshape(S, 4),
                          "moments": f[:nm].reshape(S, abi.N_METRICS, 2), "extrema": f[nm:].reshape(S, abi.N_METRICS, 2)}
            valid = self._host["counters"][:, 0]
            if (valid < self.n_total).any():
                bad = int(self.n_total - valid.min())
                from .engine import NonFiniteTrials
                raise NonFiniteTrials(f"{bad} of {self.n_total} trials generated non-finite returns or VIX <= 0 "
                                      f"(reference raises ValueError, engine.py:486-498)")
        return self._host

    @property
    def d2h_bytes(self):
        return int(self.sm.buf.numel() * 8)

    def table(self, initial_capital: float = 10000.0, exact: bool = False) -> list[dict]:
        from .summary import PERCENTILES, global_percentiles, summary_from_histograms
        h = self._fetch()
        rows = summary_from_histograms(h["hist"], self.sm.lo, self.sm.hi, h["counters"], initial_capital)
        years = self.years
        for s, (sid, row) in enumerate(zip(self.strategy_ids, rows)):
            row["id"] = sid
            row["median"] = row["p50"]
            if exact:
                qs = list(PERCENTILES)
                for q, v in zip(qs, global_percentiles(self.device_out.final_wealth[s], qs)):
                    row[f"p{q}"] = v
                row["median"] = row["p50"]
                row["dd_median"] = global_percentiles(self.device_out.max_dd[s], [50])[0]
            row["cagr"] = (row["median"] / initial_capital) ** (1 / years) - 1
            row["dd"] = row["dd_median"]
            row["trades"] = row["total_trades"] / max(row["valid"], 1) / years
            row["mean_log10_growth"] = float(h["moments"][s, 0, 0]) / max(row["valid"], 1)
        rows.sort(key=lambda r: r["median"], reverse=True)
        return rows


_ENGINES: dict = {}      # (parameter bytes, precision, device) -> PathEngine; the handle owns device scratch worth reusing
_FRAME_ENGINES: dict = {}          # single-frame engines of run_strategy_fixed (strategy.py), one per (strategy id, frame length)
_FRAME_ENGINE_CACHE_MAX = 64
_ENGINE_CACHE_MAX = 8


def _engine_for(params, keep, precision, device, cache=None):
    """Engines are pure functions of their parameter block, so repeated calls (one per horizon in
    letf.run(), letf/__init__.py:150-162) reuse the handle instead of re-creating device state.
    cache: another cache than the Monte Carlo engines' (the single-frame `run_strategy_fixed` engines hold next to no device
    memory but come one per strategy id: they get a larger cache of their own, `_FRAME_ENGINES`)"""
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
    engines, limit = (_ENGINES, _ENGINE_CACHE_MAX) if cache is None else cache
    eng = engines.get(key)
    if eng is None or eng._h is None:
        while len(engines) >= limit:
            # drop the cache entry only: earlier MonteCarloRun objects may still hold the engine for their lazy Trade_List /
            # summary calls; the handle is destroyed when the last reference goes (PathEngine.__del__)
            engines.pop(next(iter(engines)))
        eng = engines[key] = PathEngine(params, keep, precision=precision, device=dev)
    return eng


def parallel_monte_carlo_fixed(strategy_ids, time_horizon, regime_model, correlation_matrices,
                               historical_df=None, *, cfg=None, precision=None, seed=None, device=None,
                               verbose=True, summary_only=False, device_results=False):
    """Reference signature: letf/mc_runner.py:24-25.  `correlation_matrices` is accepted and unused,
    exactly as in the reference (strategy.py:55 is its only reader and ignores it).

    summary_only=True returns a `MonteCarloSummary` (device-side reduction + cross-rank exchange; no per-trial array
    leaves the device).  device_results=True returns the usual mapping with the per-trial arrays left on the device and
    copied to the host lazily, on first access."""
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
    if summary_only or device_results:
        dev = eng.run(hi - lo, first_trial=lo, seed=seed)
        run = MonteCarloRun(eng, None, strategy_ids, time_horizon, seed, lo, hi - lo, n_total, TAXABLE_IDS, device_out=dev)
        return run.summary() if summary_only else run
    # raises NonFiniteTrials like the reference's ValueError (the C-ABI stages through its own pinned buffer)
    out = eng.run_host(hi - lo, first_trial=lo, seed=seed)
    return MonteCarloRun(eng, out, strategy_ids, time_horizon, seed, lo, hi - lo, n_total, TAXABLE_IDS)
