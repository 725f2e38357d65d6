"""ctypes mirror of include/letf_b200.h and loader of the CUDA engine library.

The product library (`libletf_b200.so`, built in-tree by `__graft_entry__.build()`)
is mandatory: there is NO CPU fallback.  `load_engine_library()` raises if it is
missing, so a GPU box that lost the extension fails loudly instead of silently
computing somewhere else.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 3
N_ASSETS = 6
N_REGIMES = 2
MAX_STRATEGIES = 32
STRAT_NPARAM = 8
HIST_BINS = 4096
N_METRICS = 3

ASSET_NAMES = ("TQQQ", "UPRO", "SSO", "TMF", "SPY", "QQQ")  # cfg.ASSETS order, letf/config.py:62-135
ASSET_INDEX = {a: i for i, a in enumerate(ASSET_NAMES)}
UNDERLYING_INDEX = {"SPY": 0, "QQQ": 1, "TLT": 2}

# status codes
OK, ERR_INVALID, ERR_CUDA, ERR_NONFINITE, ERR_NOMEM = 0, 1, 2, 3, 4
# market modes
MARKET_NORMAL, MARKET_PARAM_T, MARKET_JOINT_T = 0, 1, 2
MARKET_BOOTSTRAP = 4      # legacy_hybrid block bootstrap (3 = caller-supplied series, internal)
# precision / rng
PREC_F64, PREC_F32 = 0, 1
RNG_PHILOX, RNG_TAPE, RNG_PHILOX_RECORD = 0, 1, 2
# flags
FLAG_NONFINITE_RET, FLAG_BAD_VIX, FLAG_REPAIRED, FLAG_SPELL_OVERFLOW = 1, 2, 4, 8

# strategy type enum (LETF_ST_*), keyed by cfg.STRATEGIES[*]['type'] (letf/config.py:348-437)
STRATEGY_TYPES = {
    "benchmark": 0, "sma": 1, "sma_band": 2, "portfolio": 3, "vol_targeting": 4,
    "composite": 5, "adaptive_vol": 6, "downside_vol": 7, "convex_vol": 8,
    "vol_velocity": 9, "vol_mom": 10, "skew_convex": 11, "meta_ensemble": 12,
    "regime_asymmetric": 13, "skew_kelly": 14, "mom_vol_convergence": 15,
    "conviction_compounder": 16, "benchmark_tlt": 17,
}


class AssetParams(C.Structure):
    _fields_ = [
        ("leverage", C.c_double), ("expense_ratio", C.c_double), ("borrow_spread", C.c_double),
        ("te_base", C.c_double), ("te_df", C.c_double), ("te_rho", C.c_double),
        ("te_downside", C.c_double), ("te_clip", C.c_double), ("beta", C.c_double),
        ("underlying", C.c_int32), ("beta_applies", C.c_int32),
    ]


class StrategyParams(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("asset", C.c_int32), ("asset2", C.c_int32), ("taxable", C.c_int32),
        ("weights", C.c_double * N_ASSETS), ("p", C.c_double * STRAT_NPARAM),
    ]


class JointRegime(C.Structure):
    _fields_ = [
        ("mu", C.c_double * 3), ("cov", (C.c_double * 3) * 3), ("rbar", (C.c_double * 3) * 3),
        ("nu", C.c_double), ("garch_alpha", C.c_double), ("garch_beta", C.c_double),
        ("dcc_a", C.c_double), ("dcc_b", C.c_double),
    ]


class VixRegime(C.Structure):
    _fields_ = [("phi", C.c_double), ("target", C.c_double), ("noise_std", C.c_double),
                ("jump_threshold", C.c_double), ("jump_scale", C.c_double)]


class StressRegime(C.Structure):
    _fields_ = [("liq_mu", C.c_double), ("liq_phi", C.c_double), ("liq_sigma", C.c_double),
                ("credit_mu", C.c_double), ("credit_phi", C.c_double), ("credit_sigma", C.c_double),
                ("jump_base_prob", C.c_double), ("jump_scale", C.c_double)]


class FundingModel(C.Structure):
    _fields_ = [("base", C.c_double), ("beta_vix", C.c_double), ("beta_inv_curve", C.c_double),
                ("beta_liquidity", C.c_double), ("beta_credit", C.c_double),
                ("min_spread", C.c_double), ("max_spread", C.c_double)]


class BootstrapPool(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("pad", C.c_int32),
                ("start", C.POINTER(C.c_int32)), ("block_ret", C.POINTER(C.c_double))]


class Bootstrap(C.Structure):
    _fields_ = [
        ("hist_len", C.c_int32), ("block_size", C.c_int32), ("block_min", C.c_int32), ("block_max", C.c_int32),
        ("block_mean", C.c_double), ("momentum_bias", C.c_double * 2),
        ("sampler_weight", C.c_double), ("engine_weight", C.c_double),
        ("sampler_chol", ((C.c_double * 3) * 3) * 2), ("engine_chol", ((C.c_double * 3) * 3) * 2),
        ("hist_spy", C.POINTER(C.c_double)), ("hist_qqq", C.POINTER(C.c_double)), ("hist_tlt", C.POINTER(C.c_double)),
        ("pool_a", BootstrapPool * 2), ("pool_b", BootstrapPool * 2),
    ]


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("market_mode", C.c_int32), ("sim_days", C.c_int32),
        ("buffer_days", C.c_int32), ("random_start", C.c_int32), ("antithetic_pairs", C.c_int32),
        ("moment_matching", C.c_int32), ("institutional", C.c_int32),
        ("has_vix_dynamics", C.c_int32), ("has_stress", C.c_int32), ("has_funding", C.c_int32),
        ("asset_mask", C.c_int32), ("n_strategies", C.c_int32), ("pad0", C.c_int32),
        ("sim_years", C.c_double), ("initial_capital", C.c_double),
        ("transition", (C.c_double * 2) * 2),
        ("regime_daily_mean", C.c_double * 2), ("regime_daily_std", C.c_double * 2),
        ("n_duration", C.c_int32 * 2), ("duration_samples", C.POINTER(C.c_int32) * 2),
        ("joint", JointRegime * 2),
        ("param_t_chol", ((C.c_double * 3) * 3) * 2), ("param_t_df", C.c_double),
        ("normal_mu", C.c_double * 3), ("normal_chol", (C.c_double * 3) * 3),
        ("const_vix", C.c_double), ("const_rf", C.c_double),
        ("vix", VixRegime * 2), ("vix_base", C.c_double * 2), ("irx_base", C.c_double * 2),
        ("stress", StressRegime * 2), ("funding", FundingModel),
        ("assets", AssetParams * N_ASSETS),
        ("cash_rate", C.c_double * 2), ("base_spread_bps", C.c_double * 2),
        ("rebalance_cost", C.c_double), ("apply_costs", C.c_int32), ("pad1", C.c_int32),
        ("strategies", StrategyParams * MAX_STRATEGIES),
        ("boot", Bootstrap),
        ("start_regime_prob1", C.c_double),
    ]


class Tape(C.Structure):
    _fields_ = [
        ("n_tapes", C.c_int64), ("gen_days", C.c_int32), ("sim_days", C.c_int32),
        ("max_spells", C.c_int32), ("pad", C.c_int32),
        ("start_offset", C.c_void_p), ("spells", C.c_void_p), ("x", C.c_void_p),
        ("chi2", C.c_void_p), ("vix_n", C.c_void_p), ("irx_n", C.c_void_p),
        ("liq_n", C.c_void_p), ("cred_n", C.c_void_p), ("jump_u", C.c_void_p),
        ("jump_t", C.c_void_p), ("sign", C.c_void_p), ("te_t", C.c_void_p),
        ("max_blocks", C.c_int32), ("boot_block_size", C.c_int32),
        ("boot_len", C.c_void_p), ("boot_u", C.c_void_p), ("boot_t", C.c_void_p),
        ("start_regime", C.c_void_p), ("boot_syn", C.c_void_p),
    ]


TAPE_FIELDS = ("start_offset", "spells", "x", "chi2", "vix_n", "irx_n", "liq_n", "cred_n",
               "jump_u", "jump_t", "sign", "te_t", "boot_len", "boot_u", "boot_t", "start_regime", "boot_syn")


class Outputs(C.Structure):
    _fields_ = [
        ("n_trials", C.c_int64),
        ("final_wealth", C.c_void_p), ("max_dd", C.c_void_p), ("num_trades", C.c_void_p),
        ("journal_count", C.c_void_p), ("journal_volume", C.c_void_p),
        ("start_offset", C.c_void_p), ("flags", C.c_void_p), ("regime_days", C.c_void_p),
        ("jump_days", C.c_void_p),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("hist", C.c_void_p), ("moments", C.c_void_p), ("extrema", C.c_void_p),
        ("counters", C.c_void_p),
        ("hist_lo", C.c_double * N_METRICS), ("hist_hi", C.c_double * N_METRICS),
        ("bench_strategy", C.c_int32), ("pad", C.c_int32),
    ]


class LaunchInfo(C.Structure):
    _fields_ = [
        ("grid", C.c_int32), ("block", C.c_int32), ("regs_per_thread", C.c_int32),
        ("kernel_launches", C.c_int32), ("scratch_bytes", C.c_int64),
        ("last_kernel_ms", C.c_float), ("sm_count", C.c_int32),
        ("aux_launches", C.c_int32), ("pad", C.c_int32), ("kernel", C.c_char * 48),
    ]


class Series(C.Structure):
    _fields_ = [
        ("n_paths", C.c_int64), ("sim_days", C.c_int32), ("pad", C.c_int32),
        ("asset_ret", C.c_void_p * N_ASSETS), ("tlt_ret", C.c_void_p), ("vix", C.c_void_p),
        ("regime", C.c_void_p), ("cash_ret", C.c_void_p),
    ]


ENGINE_SYMBOLS = (
    "letf_create", "letf_destroy", "letf_last_error", "letf_abi_version", "letf_run",
    "letf_run_host", "letf_summary_reset", "letf_reduce", "letf_trade_list_host", "letf_trade_lists_host",
    "letf_equity_curve_host", "letf_run_strategy_curve_host", "letf_launch_info",
    "letf_run_strategies", "letf_nearest_psd3", "letf_summary_words", "letf_summary_combine", "letf_market_returns_host",
    "letf_tax_realised_by_year", "letf_trade_lists_multi_host",
)

_HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_LIB_PATH = os.path.join(_HERE, "libletf_b200.so")
_lib = None


class EngineLibraryMissing(RuntimeError):
    pass


def load_engine_library(path: str | None = None):
    """dlopen the CUDA engine.  Raises EngineLibraryMissing if it was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or ENGINE_LIB_PATH
    if not os.path.exists(p):
        raise EngineLibraryMissing(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the path engine.")
    lib = C.CDLL(p)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    lib.letf_create.argtypes = [C.POINTER(Params), i32, i32, C.POINTER(vp)]
    lib.letf_create.restype = i32
    lib.letf_destroy.argtypes = [vp]
    lib.letf_destroy.restype = None
    lib.letf_last_error.argtypes = []
    lib.letf_last_error.restype = C.c_char_p
    lib.letf_abi_version.argtypes = []
    lib.letf_abi_version.restype = i32
    lib.letf_run.argtypes = [vp, vp, u64, i64, i64, i32, C.POINTER(Tape), C.POINTER(Outputs)]
    lib.letf_run.restype = i32
    lib.letf_run_host.argtypes = [vp, u64, i64, i64, i32, C.POINTER(Tape), C.POINTER(Outputs)]
    lib.letf_run_host.restype = i32
    lib.letf_summary_reset.argtypes = [vp, vp, C.POINTER(Summary)]
    lib.letf_summary_reset.restype = i32
    lib.letf_reduce.argtypes = [vp, vp, C.POINTER(Outputs), i64, C.POINTER(Summary)]
    lib.letf_reduce.restype = i32
    lib.letf_trade_list_host.argtypes = [vp, u64, i64, i32, i32, C.POINTER(Tape), i32, vp, vp, vp, vp, vp,
                                         C.POINTER(C.c_int32)]
    lib.letf_trade_list_host.restype = i32
    lib.letf_trade_lists_host.argtypes = [vp, u64, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.letf_trade_lists_host.restype = i32
    lib.letf_equity_curve_host.argtypes = [vp, u64, i64, i32, i32, C.POINTER(Tape), vp]
    lib.letf_equity_curve_host.restype = i32
    lib.letf_run_strategy_curve_host.argtypes = [vp, C.POINTER(Series), i32, vp, C.POINTER(C.c_int32), i32, vp, vp, vp, vp, vp,
                                                 C.POINTER(C.c_int32)]
    lib.letf_run_strategy_curve_host.restype = i32
    lib.letf_launch_info.argtypes = [vp, C.POINTER(LaunchInfo)]
    lib.letf_launch_info.restype = i32
    lib.letf_run_strategies.argtypes = [vp, vp, C.POINTER(Series), C.POINTER(Outputs)]
    lib.letf_run_strategies.restype = i32
    lib.letf_nearest_psd3.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.letf_nearest_psd3.restype = i32
    lib.letf_tax_realised_by_year.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp]
    lib.letf_tax_realised_by_year.restype = i32
    lib.letf_trade_lists_multi_host.argtypes = [vp, u64, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.letf_trade_lists_multi_host.restype = i32
    lib.letf_market_returns_host.argtypes = [vp, u64, i64, i32, C.POINTER(Tape), vp]
    lib.letf_market_returns_host.restype = i32
    lib.letf_summary_words.argtypes = [i32]
    lib.letf_summary_words.restype = i64
    lib.letf_summary_combine.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.letf_summary_combine.restype = i32
    if lib.letf_abi_version() != ABI_VERSION:
        raise EngineLibraryMissing(f"{p}: ABI version {lib.letf_abi_version()} != {ABI_VERSION}; rebuild")
    if path is None:
        _lib = lib
    return lib
