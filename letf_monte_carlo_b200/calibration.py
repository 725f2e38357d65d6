"""Drop-in for `letf.calibration.simulate_joint_returns_t` (reference: letf/calibration.py:480-590), the last inner seam of
the path: the regime-conditional multivariate Student-t generator with GARCH / DCC dynamics, run for ONE
path along a caller-supplied regime path.

`rng` is the injection point of the random variates, as in the reference: the standard normals of the correlated vector
(`rng.standard_normal((n_days, 3))`) and the chi-square mixing variables (`rng.chisquare(nu_of_the_day)`) are drawn from it
on the host, uploaded as a shock tape, and the fp64 build of the fused kernel runs the recursions on the device (variance /
correlation updates, regime blend, Cholesky factor, clipping, antithetic negation, moment matching).  The correlated
normals are `chol(cov_t) @ x` - the factor the engine uses everywhere - where NumPy's `multivariate_normal` factors with SVD:
same distribution, different variates for the same stream.
"""
from __future__ import annotations

import numpy as np

from . import abi
from . import config as _default_cfg
from .params import build_params
from .tape import HostTape

ASSETS = ("SPY_Ret", "QQQ_Ret", "TLT_Ret")


def _spells(regime_path: np.ndarray):
    """run-length encoding of a 0/1 regime path: (start regime, spell lengths)"""
    r = np.asarray(regime_path).astype(np.int64)
    if r.size and (r.min() < 0 or r.max() > 1):
        raise ValueError("regime_path values must be 0 or 1")
    cuts = np.flatnonzero(np.diff(r)) + 1
    edges = np.concatenate(([0], cuts, [r.size]))
    return int(r[0]) if r.size else 0, np.diff(edges).astype(np.int32)


def joint_case(n_days: int, regime_path, joint_model: dict, rng, antithetic: bool, cfg):
    """(params, keepalive, tape, trial id) of the one-path launch behind `simulate_joint_returns_t`"""
    regime_path = np.asarray(regime_path)[:n_days]
    if len(regime_path) != n_days:
        raise ValueError("regime_path is shorter than n_days")
    start, spells = _spells(regime_path)
    # a parameter block whose generated path IS the window: no buffer, no random offset; both regimes must be able to flip
    # (the spells alternate) and the start regime comes from the tape
    over = type("Cfg", (), {k: getattr(cfg, k) for k in dir(cfg) if k.isupper()})()
    over.USE_RANDOM_START = True
    over.RANDOM_START_METHOD = "regime_only"
    over.START_REGIME_PROBABILITIES = {0: 0.5, 1: 0.5}
    over.USE_ANTITHETIC_VARIATES = bool(antithetic)
    over.SIM_ENGINE_MODE = "institutional_v1"
    model = {"regime_params": {0: {"daily_mean": 0.0, "daily_std": 0.01}, 1: {"daily_mean": 0.0, "daily_std": 0.02}},
             "transition_matrix": np.array([[0.9, 0.1], [0.1, 0.9]]), "duration_samples": None, "vix_dynamics": None,
             "joint_return_model": joint_model, "funding_model": None, "tracking_residual_model": {}, "stress_state_model": None}
    params, keep = build_params(["S2"], n_days / 252.0, model, over)
    params.sim_days = n_days
    assert params.buffer_days == 0 and params.market_mode == abi.MARKET_JOINT_T and params.start_regime_prob1 > 0
    tape = HostTape(1, n_days, n_days, max(len(spells), 1))
    tape.spells[:len(spells), 0] = spells
    tape.start_regime[0] = start
    nu = np.where(regime_path == 0, float(joint_model["regimes"][0]["nu"]), float(joint_model["regimes"][1]["nu"]))
    tape.x[:] = np.asarray(rng.standard_normal((n_days, 3)), dtype=np.float64).reshape(tape.x.shape)
    tape.chi2[:] = np.asarray(rng.chisquare(nu), dtype=np.float64).reshape(tape.chi2.shape)
    return params, keep, tape, (1 if antithetic else 0)


def simulate_joint_returns_t(n_days: int, regime_path, joint_model: dict, rng: np.random.Generator, antithetic: bool = False, *,
                             cfg=None, device=None) -> dict:
    """-> {asset: float64[n_days]} for joint_model['assets'] (SPY_Ret, QQQ_Ret, TLT_Ret), signature of calibration.py:480."""
    from .mc_runner import _engine_for
    cfg = _default_cfg if cfg is None else cfg
    assets = list(joint_model.get("assets", ASSETS))
    n_days = int(n_days)
    if n_days == 0:
        return {a: np.zeros(0) for a in assets}
    params, keep, tape, trial = joint_case(n_days, regime_path, joint_model, rng, antithetic, cfg)
    eng = _engine_for(params, keep, "f64", device if device is not None else getattr(cfg, "GPU_DEVICE", None))
    ret = eng.market_returns(trial, rng_mode=abi.RNG_TAPE, tape=tape)
    return {a: np.ascontiguousarray(ret[:, i]) for i, a in enumerate(assets)}
