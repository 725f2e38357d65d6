"""Drop-in for `letf.simulation.engine.simulate_single_path_fixed` (reference: letf/simulation/engine.py:423-619).

One trial, reference argument tuple in, reference result dict out - for callers that drive single paths themselves
(diagnostics, notebooks).  The bulk path is `parallel_monte_carlo_fixed`; this seam launches the same fused kernel for
one trial, so it costs a whole (latency-bound, single-thread) launch per call.

Trial identity: with antithetic pairing the reference hands out even sim_ids twice, once per `antithetic` flag
(mc_runner.py:145-153); the engine's trial index is `sim_id | antithetic`.  Random numbers are the engine's Philox
streams keyed by (cfg.GPU_SEED, sim_id), not NumPy's PCG64(sim_id + 50000): distributions agree, individual paths do not.
`Regime_Path` (a T-element list nothing downstream reads) is not materialised; the regime day counts are in
`Metadata['regime_counts']`.
"""
from __future__ import annotations

from . import abi
from . import config as _default_cfg
from .params import build_params
from .trade import TAXABLE_IDS


def simulate_single_path_fixed(args, *, cfg=None, precision=None, seed=None, device=None):
    """args = (sim_id, sim_years, regime_model, correlation_matrices, strategies[, antithetic]) -> {sid: result dict}"""
    from .mc_runner import _engine_for
    cfg = _default_cfg if cfg is None else cfg
    if len(args) == 6:
        sim_id, sim_years, regime_model, _corr, strategies, antithetic = args
    else:
        sim_id, sim_years, regime_model, _corr, strategies = args
        antithetic = False
    strategies = list(strategies)
    params, keep = build_params(strategies, sim_years, regime_model, cfg)
    trial = ((int(sim_id) & ~1) | (1 if antithetic else 0)) if params.antithetic_pairs else int(sim_id)
    precision = precision or getattr(cfg, "GPU_PRECISION", "f32")
    seed = int(getattr(cfg, "GPU_SEED", 0) if seed is None else seed)
    eng = _engine_for(params, keep, precision, device if device is not None else getattr(cfg, "GPU_DEVICE", None))
    out = eng.run_host(1, first_trial=trial, seed=seed)            # raises NonFiniteTrials(ValueError) like engine.py:486-498
    fl = int(out.flags[0])
    meta = {"model_version": cfg.SIM_ENGINE_MODE, "sim_id": int(sim_id),
            "regime_counts": {0: int(out.regime_days[0, 0]), 1: int(out.regime_days[1, 0])},
            "start_method": (cfg.RANDOM_START_METHOD if getattr(cfg, "USE_RANDOM_START", False) else "default"),
            "start_offset": int(out.start_offset[0]),
            "layer_checks": {"is_valid": not (fl & abi.FLAG_REPAIRED), "repaired_after_validation": bool(fl & abi.FLAG_REPAIRED)},
            "stress_summary": {"jump_days": int(out.jump_days[0])}, "engine": "letf_monte_carlo_b200"}
    res = {}
    for s, sid in enumerate(strategies):
        n = int(out.num_trades[s, 0])
        r = {"Final_Wealth": float(out.final_wealth[s, 0]), "Num_Trades": n, "Trades_Per_Year": n / float(sim_years),
             "Max_DD": float(out.max_dd[s, 0]), "Regime_Path": None, "Metadata": meta}
        if sid in TAXABLE_IDS:
            r["Trade_List"] = eng.trade_list(trial, s, seed=seed)
        res[sid] = r
    return res
