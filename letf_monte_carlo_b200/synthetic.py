"""Synthetic (network-free) inputs of the named benchmark shapes.

`fallback_regime_model()` assembles a `regime_model` dict from the reference's own
insufficient-data fallback tables, so no market data is needed:
  joint model    letf/calibration.py:375-389
  funding model  letf/calibration.py:601-609
  stress model   letf/calibration.py:731-740
  VIX dynamics   letf/simulation/engine.py:107-111 (inline defaults)
The regime parameters / transition matrix are the ones SURVEY.md section 8c used.
`normal_market()` describes the iid-normal shapes of BASELINE.json configs 1, 2
and 5 (validation.py:33-57 and SURVEY.md section 8d).
"""
from __future__ import annotations

import copy
import types

import numpy as np

from . import config as _config


def fallback_joint_model() -> dict:
    model = {"assets": ["SPY_Ret", "QQQ_Ret", "TLT_Ret"], "regimes": {}}
    for regime in (0, 1):
        lo = regime == 0
        mu = np.array([0.08 / 252, 0.10 / 252, 0.03 / 252])
        vol = np.array([0.16, 0.24, 0.12]) if lo else np.array([0.28, 0.42, 0.16])
        c01, c02, c12 = (0.90, -0.20, -0.18) if lo else (0.96, -0.05, -0.03)
        corr = np.array([[1.0, c01, c02], [c01, 1.0, c12], [c02, c12, 1.0]])
        cov = np.outer(vol / np.sqrt(252), vol / np.sqrt(252)) * corr
        model["regimes"][regime] = {
            "mu": mu, "cov": cov, "nu": 5.0 if lo else 4.0,
            "garch_alpha": 0.06 if lo else 0.09, "garch_beta": 0.90 if lo else 0.86,
            "dcc_a": 0.02 if lo else 0.04, "dcc_b": 0.95 if lo else 0.90,
        }
    return model


def fallback_funding_model() -> dict:
    return {"base": 0.0050, "beta_vix": 0.00035, "beta_inv_curve": 0.0014, "beta_liquidity": 0.0010,
            "beta_credit": 0.0014, "min_spread": 0.0030, "max_spread": 0.0450}


def fallback_stress_model() -> dict:
    return {"regimes": {r: {
        "liq_mu": 0.10 if r == 0 else 0.35, "liq_phi": 0.90, "liq_sigma": 0.08,
        "credit_mu": 0.05 if r == 0 else 0.25, "credit_phi": 0.88, "credit_sigma": 0.07,
        "jump_base_prob": 0.0002 if r == 0 else 0.001, "jump_scale": 0.0005 if r == 0 else 0.002,
    } for r in (0, 1)}}


def fallback_regime_model(variant: str = "default") -> dict:
    """regime_model dict accepted by parallel_monte_carlo_fixed (reference and this package).

    variants (used by the golden fixtures to exercise every optional branch):
      default        joint-t + funding + stress, geometric dwell, inline VIX defaults
      no_stress      stress_state_model=None  (engine.py:503, 334)
      no_funding     funding_model=None       (engine.py:343-344: constant borrow spreads)
      durations      empirical dwell-time samples (engine.py:279-280)
      vix_te         explicit vix_dynamics + tracking_residual_model overrides incl. tanh cap
      no_joint       joint_return_model=None -> parametric Cholesky-t generator (engine.py:81-97)
      jumpy          larger jump probabilities / scales so the jump overlay fires often
    """
    m = {
        "regime_params": {0: {"daily_mean": 4e-4, "daily_std": 9.8e-3},
                          1: {"daily_mean": -3e-4, "daily_std": 2.58e-2}},
        "transition_matrix": np.array([[0.99, 0.01], [0.04, 0.96]]),
        "duration_samples": None,
        "vix_dynamics": None,
        "joint_return_model": fallback_joint_model(),
        "funding_model": fallback_funding_model(),
        "tracking_residual_model": {},
        "stress_state_model": fallback_stress_model(),
    }
    if variant == "default":
        pass
    elif variant == "no_stress":
        m["stress_state_model"] = None
    elif variant == "no_funding":
        m["funding_model"] = None
    elif variant == "durations":
        m["duration_samples"] = {0: [35, 61, 90, 120, 180, 250, 400], 1: [8, 15, 21, 30, 45, 60]}
    elif variant == "vix_te":
        m["vix_dynamics"] = {0: {"phi": 0.90, "target_vix": 16.0, "noise_std": 1.0,
                                 "jump_threshold_sigma": 1.5, "jump_scale": 6.0},
                             1: {"phi": 0.85, "target_vix": 32.0, "noise_std": 2.0,
                                 "jump_threshold_sigma": 2.5, "jump_scale": 10.0}}
        m["tracking_residual_model"] = {
            "TQQQ": {"rho": 0.45, "downside_mult": 1.6, "df": 4.0, "clip_limit": 0.004},
            "UPRO": {"rho": 0.2, "downside_mult": 1.02, "df": 6.5, "clip_limit": 0.0},
            "SSO": {"rho": 0.35, "df": 7.0},
        }
    elif variant == "no_joint":
        m["joint_return_model"] = None
    elif variant == "jumpy":
        for r in (0, 1):
            m["stress_state_model"]["regimes"][r]["jump_base_prob"] = 0.006 if r == 0 else 0.009
            m["stress_state_model"]["regimes"][r]["jump_scale"] = 0.01 if r == 0 else 0.03
    else:
        raise KeyError(variant)
    return m


def normal_market(kind: str = "c2") -> dict:
    """`synthetic_normal` market spec (iid correlated normals, fixed regime 0, constant VIX / rf).

    c1: validation.py:33-57 - one zero-drift underlying, sigma = 15 %/yr (applied to SPY, QQQ and TLT
        with unit correlation so any of the six ETF layers sees that same series)
    c2: SPY/QQQ sigma (0.16, 0.24)/sqrt(252), mu (0.08, 0.10)/252, rho 0.90;
        TLT gets the fallback low-vol parameters (0.12 vol, mu 0.03, corr -0.20/-0.18)
    """
    if kind == "c1":
        s = 0.15 / np.sqrt(252)
        chol = np.array([[s, 0, 0], [s, 0, 0], [s, 0, 0]], dtype=float)
        return {"mu": [0.0, 0.0, 0.0], "chol": chol, "vix": 15.0, "rf": 0.0}
    if kind == "c2":
        vol = np.array([0.16, 0.24, 0.12]) / np.sqrt(252)
        corr = np.array([[1.0, 0.90, -0.20], [0.90, 1.0, -0.18], [-0.20, -0.18, 1.0]])
        return {"mu": [0.08 / 252, 0.10 / 252, 0.03 / 252], "cov": np.outer(vol, vol) * corr,
                "vix": 15.0, "rf": 0.045}
    raise KeyError(kind)


def pure_leverage_assets() -> dict:
    """cfg.ASSETS with every friction zeroed: ETF return = L * underlying exactly, the arithmetic of
    the reference's vol-drag known-answer tests (validation.py:45-57, 119-133)."""
    out = copy.deepcopy(_config.ASSETS)
    for a in out.values():
        a["expense_ratio"] = 0.0
        a["borrow_spread"] = 0.0
        a["tracking_error_base"] = 0.0
    return out


def normal_regime_model(kind: str = "c2") -> dict:
    return {"regime_params": {0: {"daily_mean": 0.0, "daily_std": 0.01}, 1: {"daily_mean": 0.0, "daily_std": 0.01}},
            "transition_matrix": np.eye(2), "duration_samples": None, "vix_dynamics": None,
            "joint_return_model": None, "funding_model": None, "tracking_residual_model": {},
            "stress_state_model": None, "synthetic_normal": normal_market(kind)}


def config_namespace(**overrides):
    """A private snapshot of the config module with overrides applied (tests / bench use this
    instead of mutating module globals)."""
    ns = types.SimpleNamespace(**{k: copy.deepcopy(v) for k, v in vars(_config).items()
                                  if k.isupper()})
    for k, v in overrides.items():
        if k == "EXTRA_STRATEGIES":
            ns.STRATEGIES.update(copy.deepcopy(v))
        else:
            setattr(ns, k, v)
    return ns


# buy & hold of the series that have no strategy id in cfg.STRATEGIES
EXTRA_BENCHMARKS = {
    "B_UPRO": {"name": "UPRO Buy Hold", "type": "benchmark", "asset": "UPRO"},
    "B_QQQ": {"name": "QQQ Buy Hold", "type": "benchmark", "asset": "QQQ"},
    "B_TMF": {"name": "TMF Buy Hold", "type": "benchmark", "asset": "TMF"},
    "B_TLT": {"name": "TLT Buy Hold", "type": "benchmark", "asset": "TLT"},
}


def synthetic_history(n_days: int = 12000, seed: int = 7, start: str = "1978-01-03", nan_every: int = 0):
    """A deterministic stand-in for the historical frame the block bootstrap samples from
    (columns as letf/simulation/bootstrap.py:77-96 reads them: SPY_Ret, QQQ_Ret, TLT_Ret, VIX, IRX on a
    business-day index).  Two volatility states with persistent spells so that both regimes' pools and
    both eras (before / after the real-QQQ cutoff 1999-03-10) are populated; optional NaN holes exercise
    the nan_to_num / NaN-count rules of _create_regime_blocks."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    idx = pd.bdate_range(start=start, periods=n_days)
    state = np.zeros(n_days, dtype=int)
    t, cur = 0, 0
    while t < n_days:
        ln = int(rng.geometric(1 / (400 if cur == 0 else 120)))
        state[t:t + ln] = cur
        t += ln
        cur ^= 1
    vol = np.where(state == 0, 0.008, 0.022)
    spy = 0.0004 - 0.0008 * state + vol * rng.standard_t(6, n_days) * 0.8
    qqq = 1.2 * spy + 0.006 * rng.standard_normal(n_days)
    tlt = -0.2 * spy + 0.005 * rng.standard_normal(n_days)
    vix = np.where(state == 0, 15.0, 32.0) + 3.0 * rng.standard_normal(n_days)
    irx = np.clip(np.where(state == 0, 3.5, 1.5) + 0.5 * rng.standard_normal(n_days), 0, 15)
    if nan_every:
        spy[::nan_every] = np.nan
        vix[5::nan_every] = np.nan
        qqq[3::nan_every] = np.nan
    return pd.DataFrame({"SPY_Ret": spy, "QQQ_Ret": qqq, "TLT_Ret": tlt, "VIX": np.maximum(vix, 9.0), "IRX": irx},
                        index=idx)


def legacy_bootstrap_model(history=None, variant: str = "default") -> dict:
    """fallback regime model + a bootstrap sampler over `history` (legacy_hybrid mode, engine.py:50-80)."""
    from .bootstrap import create_bootstrap_sampler
    m = fallback_regime_model(variant)
    m["bootstrap_sampler"] = create_bootstrap_sampler(history if history is not None else synthetic_history())
    return m
