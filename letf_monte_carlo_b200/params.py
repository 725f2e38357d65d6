"""Pack the reference's parameter dicts + config globals into the POD `LetfParams`.

Input contract = what `simulate_single_path_fixed` reads from `regime_model`
(letf/simulation/engine.py:438-445) and from `cfg`.
Everything that is constant per run is hoisted here once: row-normalised
transition matrix (engine.py:264-272), `nearest_psd_matrix` of each regime's
correlation (utils.py:42-53, called every simulated day by the reference,
calibration.py:529-530), Cholesky factors of the parametric-t correlations
(engine.py:85-88).
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np

from . import abi


class UnsupportedConfig(ValueError):
    """Configuration the GPU path engine does not implement (raised, never silently ignored)."""


def nearest_psd_matrix(corr) -> np.ndarray:
    """Eigenvalue-clip projection to a unit-diagonal PSD matrix (reference: letf/utils.py:42-53)."""
    w, v = np.linalg.eigh(np.asarray(corr, dtype=float))
    w = np.where(w < 1e-8, 1e-8, w)
    m = v @ np.diag(w) @ v.T
    d = np.sqrt(np.diag(m))
    return m / np.outer(d, d)


def _normalised_transition(tm) -> np.ndarray:
    tm = np.nan_to_num(np.asarray(tm, dtype=float).copy(), nan=0.0, posinf=0.0, neginf=0.0)
    tm[tm < 0] = 0.0
    for i in range(tm.shape[0]):
        rs = tm[i].sum()
        if rs <= 0:
            tm[i, i] = 1.0
        else:
            tm[i] = tm[i] / rs
    return tm


def _set_mat(dst, src):
    src = np.asarray(src, dtype=float)
    for i in range(src.shape[0]):
        for j in range(src.shape[1]):
            dst[i][j] = float(src[i, j])


_PARAM_T_CORR = (  # letf/simulation/engine.py:85-86
    np.array([[1.0, 0.85, -0.15], [0.85, 1.0, -0.10], [-0.15, -0.10, 1.0]]),
    np.array([[1.0, 0.92, -0.45], [0.92, 1.0, -0.30], [-0.45, -0.30, 1.0]]),
)


def _underlying_of(asset: str):
    # letf/simulation/engine.py:305-317
    if asset in ("TQQQ", "QQQ"):
        return abi.UNDERLYING_INDEX["QQQ"], 1
    if asset in ("UPRO", "SSO", "SPY"):
        return abi.UNDERLYING_INDEX["SPY"], 1
    if asset == "TMF":
        return abi.UNDERLYING_INDEX["TLT"], 0
    return abi.UNDERLYING_INDEX["SPY"], 0


def pack_strategy(sid: str, conf: Mapping, taxable_ids: Sequence[str]) -> abi.StrategyParams:
    sp = abi.StrategyParams()
    kind = conf["type"]
    if kind not in abi.STRATEGY_TYPES:
        # strategy.py:1273-1277: unknown types fall through to SPY buy & hold
        kind, conf = "benchmark", {"asset": "SPY"}
    sp.type = abi.STRATEGY_TYPES[kind]
    sp.asset = abi.ASSET_INDEX.get(conf.get("asset", "SPY"), -1)
    sp.asset2 = abi.ASSET_INDEX.get(conf.get("defensive_asset", "SPY"), -1)
    sp.taxable = 1 if sid in taxable_ids else 0
    if kind == "benchmark" and conf.get("asset") == "TLT":
        sp.type = abi.STRATEGY_TYPES["benchmark_tlt"]
        sp.asset = 4
    elif sp.asset < 0 and kind != "portfolio":
        raise UnsupportedConfig(f"{sid}: unknown asset {conf.get('asset')!r}")
    g = conf.get
    p = {
        "benchmark": [],
        "sma": [g("sma_period", 200)],
        "sma_band": [g("sma_period", 200), g("band", 0.02)],
        "portfolio": [g("rebalance_freq", 21)],
        "vol_targeting": [g("target_vol"), g("lookback", 20)],
        "composite": [g("sma_period"), g("rsi_period"), g("vix_threshold")],
        "adaptive_vol": [g("bull_target"), g("bear_target"), g("lookback"), g("sma_period")],
        "downside_vol": [g("target_downside_vol"), g("lookback")],
        "convex_vol": [g("target_vol"), g("power"), g("sma_period")],
        "vol_velocity": [g("target_vol")],
        "vol_mom": [g("target_vol")],
        "skew_convex": [g("target_vol")],
        "meta_ensemble": [g("target_vol")],
        "regime_asymmetric": [g("base_target_vol"), g("crisis_target_vol"), g("vix_alarm_level"),
                              g("vol_expansion_threshold"), g("lookback_fast"), g("lookback_slow")],
        "skew_kelly": [g("base_target_vol"), g("skew_lookback"), g("vol_lookback"), g("kelly_fraction")],
        "mom_vol_convergence": [g("base_target_vol"), g("momentum_lookback"), g("vol_fast"), g("vol_slow"),
                                g("momentum_threshold")],
        "conviction_compounder": [g("base_target_vol"), g("momentum_lookback"), g("vol_lookback"),
                                  g("trend_sma"), g("rebalance_threshold")],
    }[kind]
    for i, v in enumerate(p):
        if v is None:
            raise UnsupportedConfig(f"{sid}: missing parameter #{i} for type {kind}")
        sp.p[i] = float(v)
    if kind == "portfolio":
        for a, w in conf["assets"].items():
            if a not in abi.ASSET_INDEX:
                raise UnsupportedConfig(f"{sid}: unknown asset {a!r}")
            sp.weights[abi.ASSET_INDEX[a]] = float(w)
        sp.asset = -1
    return sp


def _start_regime_prob1(method: str, historical_df, cfg) -> float:
    """P(start_regime == 1).  'regime_only': rng.choice over cfg.START_REGIME_PROBABILITIES (random_start.py:14-31).
    'historical_anchor': the anchor row is uniform on [min_days, n - 252) and its regime is VIX >= 25
    (random_start.py:91-99), i.e. Bernoulli with the share of such rows; without enough history the reference falls
    back to 'regime_only' (:188-191), and with too little for a valid range to regime 0 (:78-86)."""
    def regime_only():
        pr = cfg.START_REGIME_PROBABILITIES
        tot = float(sum(pr.values()))
        return float(pr.get(1, 0.0)) / tot
    if method == "regime_only":
        return regime_only()
    min_years = float(getattr(cfg, "MIN_HISTORY_FOR_ANCHOR", 2))
    if historical_df is None or not len(historical_df) > 252 * min_years:
        return regime_only()
    lo, hi = int(min_years * 252), len(historical_df) - 252
    if hi <= lo:
        return 0.0
    vix = np.asarray(historical_df["VIX"].values[lo:hi], dtype=float) if "VIX" in historical_df.columns \
        else np.full(hi - lo, 20.0)
    return float(np.mean(~(vix < 25)))            # regime = 0 if vix < 25 else 1 (NaN -> 1, as in the reference)


def _pack_bootstrap(p: abi.Params, keep: list, sampler, cfg) -> None:
    """bootstrap_sampler + cfg.BOOTSTRAP_* -> LetfBootstrap (host tables; letf_create copies them)."""
    from . import bootstrap as bs
    pools = bs.as_pools(sampler)
    keep.append(pools)
    b = p.boot
    b.hist_len = len(pools.hist_spy)
    b.block_size = int(pools.block_size)
    b.block_min = int(cfg.BOOTSTRAP_BLOCK_MIN); b.block_max = int(cfg.BOOTSTRAP_BLOCK_MAX)
    b.block_mean = float(cfg.BOOTSTRAP_BLOCK_MEAN)
    if not (1 <= b.block_min <= b.block_max <= b.block_size):
        # bootstrap.py:350-352 slices block[start:start+len] with len <= block_size; a longer request would
        # silently yield a short block in numpy and desynchronise the day counter
        raise UnsupportedConfig("BOOTSTRAP_BLOCK_MIN <= BOOTSTRAP_BLOCK_MAX <= BOOTSTRAP_BLOCK_SIZE is required")
    for r in range(2):
        b.momentum_bias[r] = float(cfg.BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME.get(r, 0.52))
        _set_mat(b.sampler_chol[r], np.linalg.cholesky(bs.SAMPLER_CORR[r]))
        _set_mat(b.engine_chol[r], np.linalg.cholesky(nearest_psd_matrix(bs.ENGINE_CORR[r])))
    b.sampler_weight = 0.85                                    # sample_returns' default (bootstrap.py:312)
    b.engine_weight = float(cfg.BOOTSTRAP_WEIGHT)
    dp = C.POINTER(C.c_double)
    b.hist_spy = pools.hist_spy.ctypes.data_as(dp)
    b.hist_qqq = pools.hist_qqq.ctypes.data_as(dp)
    b.hist_tlt = pools.hist_tlt.ctypes.data_as(dp)
    for name, table in (("pool_a", pools.pool_a), ("pool_b", pools.pool_b)):
        for r in range(2):
            start, ret = table[r]
            q = getattr(b, name)[r]
            q.n_blocks = len(start)      # 0: this (pool, regime) is served by synthetic Student-t blocks (bootstrap.py:214-222)
            if len(start):
                q.start = start.ctypes.data_as(C.POINTER(C.c_int32))
                q.block_ret = ret.ctypes.data_as(dp)


_NEEDS_SPY = {abi.STRATEGY_TYPES[k] for k in ("sma", "sma_band", "composite", "adaptive_vol", "convex_vol",
                                                 "meta_ensemble", "mom_vol_convergence", "conviction_compounder")}


def strategy_asset_mask(sp: abi.StrategyParams) -> int:
    """ETF series a strategy reads (SPY_Price / SPY_Ret are signals of the SMA / trend / momentum rules)."""
    t = sp.type
    m = (1 << 4) if t in _NEEDS_SPY else 0
    if t == abi.STRATEGY_TYPES["portfolio"]:
        for a in range(abi.N_ASSETS):
            if sp.weights[a] != 0.0:
                m |= 1 << a
    elif t == abi.STRATEGY_TYPES["benchmark_tlt"]:
        pass
    else:
        m |= 1 << sp.asset
        if t == abi.STRATEGY_TYPES["composite"]:
            m |= 1 << sp.asset2
    return m


def build_params(strategy_ids: Sequence[str], time_horizon: float, regime_model: Mapping, cfg,
                 *, all_assets: bool = False, taxable_ids: Sequence[str] | None = None,
                 apply_costs: bool = True):
    """-> (abi.Params, keepalive list).  `cfg` is any object with the reference's config names."""
    from .trade import TAXABLE_IDS
    taxable_ids = TAXABLE_IDS if taxable_ids is None else taxable_ids
    keep = []
    p = abi.Params()
    p.abi_version = abi.ABI_VERSION
    p.sim_days = int(time_horizon * 252)                       # engine.py:437
    p.sim_years = float(time_horizon)
    p.initial_capital = float(cfg.INITIAL_CAPITAL)
    p.apply_costs = 1 if apply_costs else 0
    if getattr(cfg, "N_REGIMES", 2) != 2:
        raise UnsupportedConfig("the path engine implements the reference's 2-regime model only")

    # ---- random start (random_start.py:132-209)
    p.buffer_days = 0
    p.random_start = 0
    if getattr(cfg, "USE_RANDOM_START", False):
        method = cfg.RANDOM_START_METHOD
        if method == "offset":
            p.buffer_days = int(cfg.RANDOM_START_BUFFER_YEARS * 252)
            p.random_start = 1 if p.buffer_days > 0 else 0
        elif method in ("regime_only", "historical_anchor"):
            # both only decide the regime the path starts in (the initial VIX they also draw is never read once the
            # VIX path is generated, engine.py:494-500); no buffer, no offset (random_start.py:169-199)
            p.start_regime_prob1 = _start_regime_prob1(method, regime_model.get("historical_df_for_anchors"), cfg)
        # any other value leaves the defaults in place, exactly as random_start.py:168-199 does
    p.antithetic_pairs = 1 if getattr(cfg, "USE_ANTITHETIC_VARIATES", False) else 0

    # ---- market generator dispatch (engine.py:45-97)
    synth = regime_model.get("synthetic_normal")
    joint = regime_model.get("joint_return_model")
    institutional = cfg.SIM_ENGINE_MODE == "institutional_v1"
    p.institutional = 1 if institutional else 0
    if synth is not None:
        p.market_mode = abi.MARKET_NORMAL
        p.antithetic_pairs = 0
    elif institutional and joint is not None:
        p.market_mode = abi.MARKET_JOINT_T
    elif regime_model.get("bootstrap_sampler") is not None and getattr(cfg, "USE_BLOCK_BOOTSTRAP", False):
        p.market_mode = abi.MARKET_BOOTSTRAP                   # engine.py:50-80
        _pack_bootstrap(p, keep, regime_model["bootstrap_sampler"], cfg)
    else:
        p.market_mode = abi.MARKET_PARAM_T
    p.moment_matching = 1 if (p.market_mode == abi.MARKET_JOINT_T and getattr(cfg, "USE_MOMENT_MATCHING", False)) else 0
    if p.market_mode != abi.MARKET_JOINT_T:
        # calibration.py:563-564: the antithetic flag only exists inside the joint-t generator
        # (twins are then identical paths, exactly as in the reference)
        pass

    # ---- regimes
    rp = regime_model.get("regime_params") or {}
    for r in range(2):
        p.regime_daily_mean[r] = float((rp.get(r) or {}).get("daily_mean", 0.0003))
        p.regime_daily_std[r] = float((rp.get(r) or {}).get("daily_std", 0.01))
    tm = regime_model.get("transition_matrix")
    tm = _normalised_transition(tm if tm is not None else np.eye(2))
    if tm.shape != (2, 2):
        raise UnsupportedConfig("transition_matrix must be 2x2")
    _set_mat(p.transition, tm)
    ds = regime_model.get("duration_samples") or {}
    for r in range(2):
        s = ds.get(r) if hasattr(ds, "get") else None
        if s is not None and len(s) > 0:
            arr = np.ascontiguousarray(np.asarray(s).astype(np.int32))
            keep.append(arr)
            p.n_duration[r] = len(arr)
            p.duration_samples[r] = arr.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            p.n_duration[r] = 0

    # ---- market models
    if p.market_mode == abi.MARKET_JOINT_T:
        if list(joint.get("assets", ["SPY_Ret", "QQQ_Ret", "TLT_Ret"])) != ["SPY_Ret", "QQQ_Ret", "TLT_Ret"]:
            raise UnsupportedConfig("joint_return_model['assets'] must be SPY_Ret, QQQ_Ret, TLT_Ret")
        if len(joint["regimes"]) != 2:
            raise UnsupportedConfig("joint_return_model must have 2 regimes")
        for r in range(2):
            jr = joint["regimes"][r]
            cov = np.asarray(jr["cov"], dtype=float)
            for a in range(3):
                p.joint[r].mu[a] = float(np.asarray(jr["mu"], dtype=float)[a])
            _set_mat(p.joint[r].cov, cov)
            sd = np.sqrt(np.diag(cov))
            _set_mat(p.joint[r].rbar, nearest_psd_matrix(cov / np.outer(sd, sd)))  # calibration.py:529-530
            p.joint[r].nu = float(jr["nu"])
            p.joint[r].garch_alpha = float(jr.get("garch_alpha", 0.06))
            p.joint[r].garch_beta = float(jr.get("garch_beta", 0.90))
            p.joint[r].dcc_a = float(jr.get("dcc_a", 0.02))
            p.joint[r].dcc_b = float(jr.get("dcc_b", 0.95))
            if not p.joint[r].nu >= 2.0:
                raise UnsupportedConfig("joint model nu must be >= 2")
    for r in range(2):
        _set_mat(p.param_t_chol[r], np.linalg.cholesky(nearest_psd_matrix(_PARAM_T_CORR[r])))
    p.param_t_df = float(cfg.STUDENT_T_DF)
    if synth is not None:
        mu = np.asarray(synth.get("mu", [0.0, 0.0, 0.0]), dtype=float)
        if "chol" in synth:
            ch = np.asarray(synth["chol"], dtype=float)
        else:
            ch = np.linalg.cholesky(np.asarray(synth["cov"], dtype=float))
        for a in range(3):
            p.normal_mu[a] = float(mu[a])
        _set_mat(p.normal_chol, ch)
        p.const_vix = float(synth.get("vix", 15.0))
        p.const_rf = float(synth.get("rf", 0.045))       # engine.py:518, regime 0
    p.has_vix_dynamics = 0 if synth is not None else 1

    vd = regime_model.get("vix_dynamics") or {}
    base = {0: 15.0, 1: 35.0}                             # engine.py:101
    for r in range(2):
        q = vd.get(r, {}) or {}
        p.vix[r].phi = float(q.get("phi", 0.88))
        p.vix[r].target = float(q.get("target_vix", base[r]))
        p.vix[r].noise_std = float(q.get("noise_std", 1.2))
        p.vix[r].jump_threshold = float(q.get("jump_threshold_sigma", 2.0))
        p.vix[r].jump_scale = float(q.get("jump_scale", 8.0))
        p.vix_base[r] = base[r]
        p.irx_base[r] = 3.5 if r == 0 else 1.5           # engine.py:121

    sm = regime_model.get("stress_state_model")
    p.has_stress = 1 if (institutional and sm is not None and synth is None) else 0
    if sm is not None:
        for r in range(2):
            q = sm["regimes"].get(r, {}) or {}
            s = p.stress[r]
            s.liq_mu = float(q.get("liq_mu", 0.1)); s.liq_phi = float(q.get("liq_phi", 0.9))
            s.liq_sigma = float(q.get("liq_sigma", 0.08))
            s.credit_mu = float(q.get("credit_mu", 0.05)); s.credit_phi = float(q.get("credit_phi", 0.88))
            s.credit_sigma = float(q.get("credit_sigma", 0.07))
            s.jump_base_prob = float(q.get("jump_base_prob", 0.001))
            s.jump_scale = float(q.get("jump_scale", 0.005))
    fm = regime_model.get("funding_model")
    p.has_funding = 1 if (institutional and fm is not None) else 0
    if fm is not None:
        f = p.funding
        f.base = float(fm["base"]); f.beta_vix = float(fm["beta_vix"])
        f.beta_inv_curve = float(fm["beta_inv_curve"])
        f.beta_liquidity = float(fm.get("beta_liquidity", 0.0012))
        f.beta_credit = float(fm.get("beta_credit", 0.0018))
        f.min_spread = float(fm["min_spread"]); f.max_spread = float(fm["max_spread"])

    # ---- asset layers (engine.py:360-420)
    trm = regime_model.get("tracking_residual_model") or {}
    for name, i in abi.ASSET_INDEX.items():
        conf = cfg.LETF_CONFIGS[name]
        te = trm.get(name, {}) or {}
        a = p.assets[i]
        a.leverage = float(conf["leverage"]); a.expense_ratio = float(conf["expense_ratio"])
        a.borrow_spread = float(conf.get("borrow_spread", 0.0))
        a.te_base = float(conf["tracking_error_base"])
        a.te_df = float(te.get("df", conf["tracking_error_df"]))
        a.te_rho = float(te.get("rho", 0.3))
        a.te_downside = float(min(te.get("downside_mult", 1.05), 1.10))
        cl = te.get("clip_limit", None)
        a.te_clip = float(cl) if (cl is not None and cl > 0) else 0.0
        a.beta = float(conf.get("beta_to_spy", 1.0))
        a.underlying, a.beta_applies = _underlying_of(name)

    # ---- strategies
    if len(strategy_ids) > abi.MAX_STRATEGIES:
        raise UnsupportedConfig(f"at most {abi.MAX_STRATEGIES} strategies per launch")
    p.n_strategies = len(strategy_ids)
    mask = 0
    for k, sid in enumerate(strategy_ids):
        sp = pack_strategy(sid, cfg.STRATEGIES[sid], taxable_ids)
        p.strategies[k] = sp
        mask |= strategy_asset_mask(sp)
    p.asset_mask = 0x3F if all_assets else mask

    for r in range(2):
        p.cash_rate[r] = float(cfg.CASH_RATE_BY_REGIME[r])
        p.base_spread_bps[r] = float(cfg.BASE_SPREAD_BPS[r])
    p.rebalance_cost = float(cfg.REBALANCE_COST_PER_DOLLAR)
    return p, keep


def gen_days(p: abi.Params) -> int:
    return int(p.sim_days + p.buffer_days)
