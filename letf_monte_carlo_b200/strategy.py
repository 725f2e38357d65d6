"""Drop-in for `letf.strategy.run_strategy_fixed` (reference: letf/strategy.py:55-1277) on the GPU.

The reference runs one strategy rule over one simulated (or historical) frame and returns its daily equity curve and
trade count, logging share trades into a `TradeJournal` for the taxable ids.  Here the frame's return columns are
uploaded, the fused kernel's strategy stage runs on that single path in its fp64 parity build with the equity-curve and
journal logs switched on, and the same two values come back.

Prices are rebuilt from the `<asset>_Ret` columns exactly as the engine builds them (`100 * cumprod(1 + r)`,
letf/simulation/engine.py:546-547); a frame whose `_Price` columns are not the cumulative product of its `_Ret` columns
(e.g. raw historical prices) would be read differently by the reference's SMA / trend rules.
"""
from __future__ import annotations

import numpy as np

from . import abi
from . import config as _default_cfg
from .params import build_params
from .trade import TAXABLE_IDS, Trade


def run_strategy_fixed(df, strategy_id, regime_path, correlation_matrices=None, apply_costs=True, trade_journal=None, *,
                       cfg=None, precision="f64", device=None):
    """-> (equity_curve: pandas.Series[T], num_trades: int); signature of strategy.py:55 (`correlation_matrices` is
    accepted and unused, as in the reference)."""
    import pandas as pd
    from .mc_runner import _FRAME_ENGINE_CACHE_MAX, _FRAME_ENGINES, _engine_for
    from .utils import regime_path_for_frame
    cfg = _default_cfg if cfg is None else cfg
    T = len(df)
    # strategy.py:64-78: a missing regime path, or one whose length does not match the frame, is inferred from the frame
    regime_path = regime_path_for_frame(df, regime_path)
    model = {"regime_params": {}, "transition_matrix": np.eye(2), "joint_return_model": None}
    taxable = TAXABLE_IDS if trade_journal is not None else ()
    params, keep = build_params([strategy_id], max(T, 1) / 252.0, model, cfg, taxable_ids=taxable, apply_costs=bool(apply_costs))
    frame = {a: df[f"{a}_Ret"].fillna(0).to_numpy() for a in abi.ASSET_NAMES if f"{a}_Ret" in df.columns}
    if "SPY" not in frame:
        raise KeyError("run_strategy_fixed needs SPY_Ret (the trend / momentum signals read SPY_Price)")
    eng = _engine_for(params, keep, precision, device if device is not None else getattr(cfg, "GPU_DEVICE", None),
                      cache=(_FRAME_ENGINES, _FRAME_ENGINE_CACHE_MAX))
    curve, n_trades, trades = eng.run_strategy_curve(
        frame, 0,
        tlt_ret=df["TLT_Ret"].fillna(0).to_numpy() if "TLT_Ret" in df.columns else None,
        vix=df["VIX"].to_numpy() if "VIX" in df.columns else None,
        cash_ret=df["Cash_Ret"].to_numpy() if "Cash_Ret" in df.columns else None,      # the reference reads the column (strategy.py:138)
        regime=np.asarray(regime_path, dtype=np.int32))
    if trade_journal is not None:
        for t in trades:
            trade_journal.trades.append(Trade(t["day_index"], t["asset"], t["action"], t["shares"], t["price"], t["dollar_amount"]))
            sign = 1.0 if t["action"] == "BUY" else -1.0
            trade_journal.positions[t["asset"]] = trade_journal.positions.get(t["asset"], 0.0) + sign * t["shares"]
    return pd.Series(curve, index=df.index), n_trades
