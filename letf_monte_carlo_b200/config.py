"""Run configuration read by the path engine.

Same module-level names (and default values) as the reference's `letf/config.py`
for everything the Monte Carlo path reads, so that user code
such as ``cfg.NUM_SIMULATIONS = 10`` keeps working.  Everything the path does not
read (cache paths, date pickers, plotting) is deliberately absent.

Unlike the reference - whose loky workers re-import the module and silently lose
parent-side mutations (SURVEY.md 8b "Threading") - the engine snapshots these
globals in the calling process at launch time.
"""
import multiprocessing as _mp

INITIAL_CAPITAL = 10000                      # letf/config.py:14
TIME_HORIZONS = [1, 2, 5, 10, 20, 30, 40, 50]


def _asset(lev, er, und, beta, te, df, spread):
    return {"leverage": lev, "expense_ratio": er, "underlying": und, "beta_to_spy": beta,
            "tracking_error_base": te, "tracking_error_df": df, "borrow_spread": spread}


# letf/config.py:62-135 (only the keys the path reads)
ASSETS = {
    "TQQQ": _asset(3.0, 0.0086, "QQQ", 1.0, 0.0002, 5, 0.0075),
    "UPRO": _asset(3.0, 0.0091, "SPY", 1.0, 0.00015, 5, 0.0060),
    "SSO": _asset(2.0, 0.0089, "SPY", 1.0, 0.0001, 5, 0.0050),
    "TMF": _asset(3.0, 0.0108, "TLT", -0.3, 0.0003, 5, 0.0040),
    "SPY": _asset(1.0, 0.000945, "SPY", 1.0, 0.00005, 10, 0.0),
    "QQQ": _asset(1.0, 0.0020, "QQQ", 1.0, 0.00005, 10, 0.0),
}
LETF_CONFIGS = ASSETS

BASE_SPREAD_BPS = {0: 2, 1: 8}               # letf/config.py:141
REBALANCE_COST_PER_DOLLAR = 0.0001
CASH_RATE_BY_REGIME = {0: 0.010, 1: -0.020}  # letf/config.py:145-148

N_WORKERS = max(1, _mp.cpu_count() - 2)      # kept for API compatibility; unused on the GPU
NUM_SIMULATIONS = 200
USE_ANTITHETIC_VARIATES = True
USE_MOMENT_MATCHING = True
USE_LATIN_HYPERCUBE = False
DEBUG = False
USE_GPU = True                               # the reference's own dispatch flag (mc_runner.py:57)

N_REGIMES = 2
REGIME_NAMES = {0: "Low Vol", 1: "High Vol"}

STUDENT_T_DF = 5                             # letf/config.py:239
USE_BLOCK_BOOTSTRAP = True
BOOTSTRAP_WEIGHT = 0.80
BOOTSTRAP_BLOCK_MIN = 21                     # letf/config.py:221-224
BOOTSTRAP_BLOCK_MAX = 168
BOOTSTRAP_BLOCK_MEAN = 84
BOOTSTRAP_BLOCK_SIZE = BOOTSTRAP_BLOCK_MAX
BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME = {0: 0.54, 1: 0.505}   # letf/config.py:227-230
SIM_ENGINE_MODE = "institutional_v1"         # letf/config.py:267

USE_RANDOM_START = True                      # letf/config.py:308-345
RANDOM_START_METHOD = "offset"
START_REGIME_PROBABILITIES = {0: 0.80, 1: 0.20}
RANDOM_START_BUFFER_YEARS = 5
MIN_HISTORY_FOR_ANCHOR = 2
RANDOMIZE_INITIAL_VIX = True
INITIAL_VIX_RANGE = {0: (12, 20), 1: (25, 45)}

# letf/config.py:348-437
STRATEGIES = {
    "S1": {"name": "TQQQ Buy Hold", "type": "benchmark", "asset": "TQQQ"},
    "S2": {"name": "SPY Buy Hold", "type": "benchmark", "asset": "SPY"},
    "S3": {"name": "SSO BuyHold (2x)", "type": "benchmark", "asset": "SSO"},
    "S4": {"name": "200-SMA Simple", "type": "sma", "asset": "TQQQ", "sma_period": 200},
    "S5": {"name": "SMA +/-2% Band", "type": "sma_band", "asset": "TQQQ", "sma_period": 200, "band": 0.02},
    "S6": {"name": "60/40 TQQQ/TMF", "type": "portfolio", "assets": {"TQQQ": 0.6, "TMF": 0.4},
           "rebalance_freq": 21},
    "S7": {"name": "Vol Targeting (20%)", "type": "vol_targeting", "asset": "TQQQ", "target_vol": 0.20,
           "lookback": 20},
    "S8": {"name": "Composite Regime", "type": "composite", "asset": "TQQQ", "defensive_asset": "SPY",
           "sma_period": 200, "rsi_period": 14, "vix_threshold": 25.0},
    "S9": {"name": "Adaptive Vol Target", "type": "adaptive_vol", "asset": "TQQQ", "bull_target": 0.35,
           "bear_target": 0.12, "lookback": 20, "sma_period": 200},
    "S10": {"name": "Sortino Optimize", "type": "downside_vol", "asset": "TQQQ",
            "target_downside_vol": 0.15, "lookback": 20},
    "S11": {"name": "Hyper-Convex", "type": "convex_vol", "asset": "TQQQ", "target_vol": 0.25,
            "power": 1.2, "sma_period": 200},
    "S12": {"name": "Vol-Velocity", "type": "vol_velocity", "asset": "TQQQ", "target_vol": 0.22},
    "S13": {"name": "VoV Momentum", "type": "vol_mom", "asset": "TQQQ", "target_vol": 0.25},
    "S14": {"name": "Skewness-Adjusted", "type": "skew_convex", "asset": "TQQQ", "target_vol": 0.25},
    "S15": {"name": "Meta-Ensemble", "type": "meta_ensemble", "asset": "TQQQ", "target_vol": 0.28},
    "S16": {"name": "Crisis Alpha", "type": "regime_asymmetric", "asset": "TQQQ", "base_target_vol": 0.30,
            "crisis_target_vol": 0.08, "vix_alarm_level": 25, "vol_expansion_threshold": 1.5,
            "lookback_fast": 5, "lookback_slow": 60},
    "S17": {"name": "Tail Risk Optimizer", "type": "skew_kelly", "asset": "TQQQ", "base_target_vol": 0.30,
            "skew_lookback": 60, "vol_lookback": 20, "kelly_fraction": 0.7},
    "S18": {"name": "Mom. Vol Conv.", "type": "mom_vol_convergence", "asset": "TQQQ",
            "base_target_vol": 0.28, "momentum_lookback": 126, "vol_fast": 10, "vol_slow": 60,
            "momentum_threshold": 0.05},
    "S19": {"name": "Conviction Compounder", "type": "conviction_compounder", "asset": "TQQQ",
            "base_target_vol": 0.32, "momentum_lookback": 126, "vol_lookback": 20, "trend_sma": 100,
            "rebalance_threshold": 0.05},
}

# ---- engine-side knobs that have no counterpart in the reference -------------------------
GPU_PRECISION = "f32"       # "f32" throughput mode | "f64" parity mode
GPU_SEED = 20240229          # Philox key; results are a pure function of (seed, trial id)
GPU_DEVICE = None            # None -> LOCAL_RANK or 0
