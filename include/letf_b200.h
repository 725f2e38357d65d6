/*
 * letf_b200.h — C-ABI of the B200-native LETF Monte Carlo path engine.
 *
 * This is the drop-in boundary for the one hot path of djbm10/LETF-Monte-Carlo:
 *   letf/mc_runner.py:24-210        parallel_monte_carlo_fixed   (trial fan-out)
 *   letf/simulation/engine.py:423   simulate_single_path_fixed   (one trial)
 *   letf/strategy.py:55             run_strategy_fixed           (per-step rule)
 * The reference has no FFI; its seam is "plain parameter dicts in, per-trial
 * result dicts out".  Every entry point below takes POD structs, raw pointers
 * and sizes only (no torch / numpy types), returns an int status (0 = ok) and
 * never throws across the boundary.  The Python binding a maintainer of the
 * reference would add is a ctypes stub; see INTEGRATION.md.
 *
 * Conventions
 *   - assets   are indexed in cfg.ASSETS order (letf/config.py:62-135):
 *              0 TQQQ, 1 UPRO, 2 SSO, 3 TMF, 4 SPY, 5 QQQ
 *   - underlyings: 0 SPY, 1 QQQ, 2 TLT  (joint model order, calibration.py:368)
 *   - regimes  0 low-vol, 1 high-vol (cfg.N_REGIMES == 2, config.py:169)
 *   - "sim day" i in [0,T) is a day of the simulated window, "generated day"
 *     g in [0,G), G = T + buffer_days, i = g - start_offset (engine.py:457-483)
 *   - trial k maps to the reference's sim_id: with antithetic pairing
 *     (mc_runner.py:145-153) sim_id = k & ~1 and antithetic = k & 1,
 *     otherwise sim_id = k.
 */
#ifndef LETF_B200_H
#define LETF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LETF_ABI_VERSION 3

#define LETF_N_ASSETS 6
#define LETF_N_UNDERLYINGS 3
#define LETF_N_REGIMES 2
#define LETF_MAX_STRATEGIES 32
#define LETF_STRAT_NPARAM 8

/* status codes */
enum {
  LETF_OK = 0,
  LETF_ERR_INVALID = 1,      /* bad argument / unsupported configuration */
  LETF_ERR_CUDA = 2,         /* CUDA runtime error (see letf_last_error) */
  LETF_ERR_NONFINITE = 3,    /* a trial produced non-finite returns / VIX <= 0
                                (reference raises ValueError, engine.py:486-498) */
  LETF_ERR_NOMEM = 4
};

/* market generator (engine.py:45-97) */
enum {
  LETF_MARKET_NORMAL = 0,   /* iid correlated normals, fixed regime 0, constant VIX / rf
                               (validation.py:45-57 and the config-2/5 synthetic shapes) */
  LETF_MARKET_PARAM_T = 1,  /* regime-switching Cholesky Student-t, engine.py:81-97 */
  LETF_MARKET_JOINT_T = 2,  /* institutional_v1 multivariate-t GARCH/DCC, calibration.py:480-590 */
  /* 3 is reserved (caller-supplied series, letf_run_strategies) */
  LETF_MARKET_BOOTSTRAP = 4 /* legacy_hybrid two-pool block bootstrap blended with correlated Student-t noise,
                               bootstrap.py:310-458 + engine.py:50-80 */
};

/* strategy types (cfg.STRATEGIES[*]['type'], config.py:348-437) */
enum {
  LETF_ST_BENCHMARK = 0,        /* strategy.py:85-95   */
  LETF_ST_SMA = 1,              /* strategy.py:98-152  */
  LETF_ST_SMA_BAND = 2,
  LETF_ST_PORTFOLIO = 3,        /* strategy.py:155-241 */
  LETF_ST_VOL_TARGETING = 4,    /* strategy.py:244-303 */
  LETF_ST_COMPOSITE = 5,        /* strategy.py:308-387 */
  LETF_ST_ADAPTIVE_VOL = 6,     /* strategy.py:391-468 */
  LETF_ST_DOWNSIDE_VOL = 7,     /* strategy.py:472-534 */
  LETF_ST_CONVEX_VOL = 8,       /* strategy.py:539-599 */
  LETF_ST_VOL_VELOCITY = 9,     /* strategy.py:603-657 */
  LETF_ST_VOL_MOM = 10,         /* strategy.py:661-719 */
  LETF_ST_SKEW_CONVEX = 11,     /* strategy.py:724-783 */
  LETF_ST_META_ENSEMBLE = 12,   /* strategy.py:788-858 */
  LETF_ST_REGIME_ASYMMETRIC = 13, /* strategy.py:863-936 */
  LETF_ST_SKEW_KELLY = 14,      /* strategy.py:941-1042 */
  LETF_ST_MOM_VOL_CONV = 15,    /* strategy.py:1047-1144 */
  LETF_ST_CONVICTION = 16,      /* strategy.py:1149-1271 */
  LETF_ST_BENCHMARK_TLT = 17,   /* buy&hold of the raw TLT_Ret column (engine.py:548) */
  LETF_ST__COUNT = 18
};

/* precision of the fused kernel */
enum {
  LETF_PREC_F64 = 0,  /* parity mode: every state variable fp64 (reference arithmetic) */
  LETF_PREC_F32 = 1   /* throughput mode: fp32 shocks/returns/signals, fp64 wealth compounding */
};

/* where the shocks come from */
enum {
  LETF_RNG_PHILOX = 0,         /* on-device Philox4x32-10 keyed by (seed, trial, day, slot) */
  LETF_RNG_TAPE = 1,           /* pre-drawn variates read from a LetfTape (bit/1e-6 parity mode) */
  LETF_RNG_PHILOX_RECORD = 2   /* Philox, and every variate consumed is also written to a LetfTape */
};

/* per-trial status flags (engine.py:486-498, 206-231, 552-562) */
enum {
  LETF_FLAG_NONFINITE_RET = 1u << 0,  /* reference raises ValueError */
  LETF_FLAG_BAD_VIX = 1u << 1,        /* VIX non-finite or <= 0: reference raises */
  LETF_FLAG_REPAIRED = 1u << 2,       /* validate_simulation_layers failed -> 'repaired_after_validation' */
  LETF_FLAG_SPELL_OVERFLOW = 1u << 3  /* tape ran out of dwell-time draws */
};

typedef struct LetfAssetParams {
  double leverage;        /* cfg.ASSETS[a]['leverage'] */
  double expense_ratio;   /* annual */
  double borrow_spread;   /* fallback spread when no funding model (engine.py:344) */
  double te_base;         /* tracking_error_base */
  double te_df;           /* tracking_residual_model[a]['df'] or tracking_error_df */
  double te_rho;          /* default 0.3 */
  double te_downside;     /* min(downside_mult, 1.10), engine.py:396 */
  double te_clip;         /* <= 0: no tanh cap (engine.py:198-203) */
  double beta;            /* beta_to_spy; ignored for TMF (engine.py:305-317) */
  int32_t underlying;     /* 0 SPY, 1 QQQ, 2 TLT */
  int32_t beta_applies;   /* 0 for TMF */
} LetfAssetParams;

typedef struct LetfStrategyParams {
  int32_t type;           /* LETF_ST_* */
  int32_t asset;          /* main asset index (benchmark / traded asset) */
  int32_t asset2;         /* composite: defensive asset */
  int32_t taxable;        /* id in TAXABLE_IDS: journal share-trades (trade.py:94) */
  double weights[LETF_N_ASSETS];     /* portfolio target weights (0 = not held) */
  double p[LETF_STRAT_NPARAM];       /* type-specific numbers, see DESIGN.md table */
} LetfStrategyParams;

typedef struct LetfJointRegime {      /* joint_return_model['regimes'][r] */
  double mu[3];
  double cov[3][3];
  double rbar[3][3];      /* nearest_psd_matrix(corr(cov)) (utils.py:42-53), host-computed once */
  double nu, garch_alpha, garch_beta, dcc_a, dcc_b;
} LetfJointRegime;

typedef struct LetfVixRegime {        /* vix_dynamics[r], defaults engine.py:107-111 */
  double phi, target, noise_std, jump_threshold, jump_scale;
} LetfVixRegime;

typedef struct LetfStressRegime {     /* stress_state_model['regimes'][r], calibration.py:731-740 */
  double liq_mu, liq_phi, liq_sigma;
  double credit_mu, credit_phi, credit_sigma;
  double jump_base_prob, jump_scale;
} LetfStressRegime;

typedef struct LetfFundingModel {     /* funding_model, calibration.py:601-609 */
  double base, beta_vix, beta_inv_curve, beta_liquidity, beta_credit, min_spread, max_spread;
} LetfFundingModel;

typedef struct LetfBootstrapPool {     /* one block pool of one regime, bootstrap.py:49-160 */
  int32_t n_blocks;
  int32_t pad;
  const int32_t* start;       /* HOST [n_blocks] first history row of each block (copied by letf_create) */
  const double* block_ret;    /* HOST [n_blocks] cumulative SPY return of the block: the selection key */
} LetfBootstrapPool;

typedef struct LetfBootstrap {         /* bootstrap_sampler + cfg.BOOTSTRAP_* (legacy_hybrid only) */
  int32_t hist_len;           /* rows of the cleaned history */
  int32_t block_size;         /* cfg.BOOTSTRAP_BLOCK_SIZE: rows per pool block (168) */
  int32_t block_min, block_max;  /* cfg.BOOTSTRAP_BLOCK_MIN / MAX, bootstrap.py:302-308 */
  double block_mean;          /* cfg.BOOTSTRAP_BLOCK_MEAN: geometric p = 1 / mean */
  double momentum_bias[2];    /* cfg.BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME */
  double sampler_weight;      /* sample_returns(bootstrap_weight=0.85), bootstrap.py:312 */
  double engine_weight;       /* cfg.BOOTSTRAP_WEIGHT, engine.py:64, 77-79 */
  double sampler_chol[2][3][3];  /* chol(corr_low_vol / corr_high_vol), bootstrap.py:404-417 */
  double engine_chol[2][3][3];   /* chol(nearest_psd(corr_low / corr_high)), engine.py:60-63 */
  const double* hist_spy;     /* HOST [hist_len] SPY_Ret, NaN -> 0 (bootstrap.py:131) */
  const double* hist_qqq;     /* HOST [hist_len] QQQ_Ret, NaN -> 0 */
  const double* hist_tlt;     /* HOST [hist_len] TLT_Ret, NaN -> 0 */
  /* a pool may be EMPTY (n_blocks 0): its blocks are then synthetic Student-t returns, bootstrap.py:214-222, 273-300 */
  LetfBootstrapPool pool_a[2];   /* "economy" pool: SPY comes from these blocks */
  LetfBootstrapPool pool_b[2];   /* "tech" pool (real-QQQ era): QQQ and TLT come from these blocks */
} LetfBootstrap;

typedef struct LetfParams {
  int32_t abi_version;         /* LETF_ABI_VERSION */
  int32_t market_mode;         /* LETF_MARKET_* */
  int32_t sim_days;            /* T = int(years*252), engine.py:437 */
  int32_t buffer_days;         /* int(RANDOM_START_BUFFER_YEARS*252) or 0, random_start.py:179 */
  int32_t random_start;        /* 1: start_offset ~ integers(min(50,buf//4), buf), random_start.py:55-57 */
  int32_t antithetic_pairs;    /* cfg.USE_ANTITHETIC_VARIATES pairing, mc_runner.py:145-153 */
  int32_t moment_matching;     /* cfg.USE_MOMENT_MATCHING, calibration.py:579-588 (JOINT_T only) */
  int32_t institutional;       /* cfg.SIM_ENGINE_MODE == 'institutional_v1' (gates funding + stress) */
  int32_t has_vix_dynamics;    /* generated VIX path (always, except MARKET_NORMAL) */
  int32_t has_stress;          /* stress_state_model given and institutional (engine.py:503) */
  int32_t has_funding;         /* funding_model given and institutional (engine.py:328) */
  int32_t asset_mask;          /* which of the 6 ETF series must be produced */
  int32_t n_strategies;
  int32_t pad0;
  double sim_years;            /* time_horizon, for Trades_Per_Year (engine.py:590) */
  double initial_capital;      /* cfg.INITIAL_CAPITAL */

  /* regimes (engine.py:255-302) */
  double transition[2][2];     /* row-normalised copy */
  double regime_daily_mean[2]; /* regime_params[r]['daily_mean'] */
  double regime_daily_std[2];
  int32_t n_duration[2];       /* len(duration_samples[r]) or 0 -> geometric dwell */
  const int32_t* duration_samples[2]; /* HOST pointers, copied by letf_create */

  /* market models */
  LetfJointRegime joint[2];
  double param_t_chol[2][3][3];  /* chol(nearest_psd(corr_r)), engine.py:85-88 */
  double param_t_df;             /* cfg.STUDENT_T_DF */
  double normal_mu[3];           /* MARKET_NORMAL: daily mean of SPY,QQQ,TLT */
  double normal_chol[3][3];      /* MARKET_NORMAL: Cholesky factor of the daily covariance */
  double const_vix;              /* MARKET_NORMAL: VIX level (15) */
  double const_rf;               /* MARKET_NORMAL: annual risk-free (0.045, engine.py:518) */

  LetfVixRegime vix[2];
  double vix_base[2];            /* {15, 35} engine.py:101 */
  double irx_base[2];            /* {3.5, 1.5} engine.py:121 */
  LetfStressRegime stress[2];
  LetfFundingModel funding;

  LetfAssetParams assets[LETF_N_ASSETS];
  double cash_rate[2];           /* cfg.CASH_RATE_BY_REGIME */
  double base_spread_bps[2];     /* cfg.BASE_SPREAD_BPS */
  double rebalance_cost;         /* cfg.REBALANCE_COST_PER_DOLLAR */
  int32_t apply_costs;           /* run_strategy_fixed(apply_costs=True) */
  int32_t pad1;

  LetfStrategyParams strategies[LETF_MAX_STRATEGIES];

  LetfBootstrap boot;            /* MARKET_BOOTSTRAP only */

  double start_regime_prob1;     /* P(trial starts in regime 1): cfg.START_REGIME_PROBABILITIES for RANDOM_START_METHOD
                                    'regime_only' (random_start.py:14-31), the share of valid anchor rows with VIX >= 25
                                    for 'historical_anchor' (:60-129, the anchor row is uniform); 0 = regime 0 (default,
                                    and always with 'offset') */
} LetfParams;

/*
 * Pre-drawn unit-scale variates, one column per tape index.  All arrays are
 * day-major ("[day][n_tapes]") so that a warp reads contiguous memory.
 * tape index of trial k = antithetic_pairs ? k >> 1 : k  (twins share their draws
 * exactly as the reference's twins share sim_id, calibration.py:563-564).
 * Draw ledger: SURVEY.md Appendix B.
 */
typedef struct LetfTape {
  int64_t n_tapes;
  int32_t gen_days;              /* G */
  int32_t sim_days;              /* T */
  int32_t max_spells;
  int32_t pad;
  int32_t* start_offset;         /* [n]            random_start.py:57 */
  int32_t* spells;               /* [max_spells][n] engine.py:280/283 (already >= 1) */
  double* x;                     /* [G][3][n] N(0,1) (JOINT_T, NORMAL) or t(df) (PARAM_T, BOOTSTRAP: engine.py:68) */
  double* chi2;                  /* [G][n]   chisquare(nu_regime)      calibration.py:565 */
  double* vix_n;                 /* [G][n]   N(0,1)                    engine.py:115 */
  double* irx_n;                 /* [G][n]   N(0,1)                    engine.py:122 */
  double* liq_n;                 /* [T][n]   N(0,1)                    calibration.py:814 */
  double* cred_n;                /* [T][n]   N(0,1)                    calibration.py:815 */
  double* jump_u;                /* [T][n]   U[0,1)                    calibration.py:827 */
  double* jump_t;                /* [T][n]   t(5)                      calibration.py:828 */
  double* sign;                  /* [T][n]   -1 / +1                   engine.py:509 */
  double* te_t;                  /* [6][T][n] t(te_df[a])              engine.py:185 */
  /* MARKET_BOOTSTRAP only (NULL otherwise); block k of a trial, at most max_blocks = G / block_min + 2 */
  int32_t max_blocks;
  int32_t boot_block_size;       /* rows per synthetic block in boot_syn (= LetfBootstrap.block_size), 0 without boot_syn */
  int32_t* boot_len;             /* [max_blocks][n]    geometric(1 / block_mean)            bootstrap.py:304 */
  double* boot_u;                /* [max_blocks][6][n] U[0,1): momentum A, block A, offset A, momentum B,
                                    block B (uniform index or weighted choice), offset B      bootstrap.py:212-262, 348-374 */
  double* boot_t;                /* [G][3][n] t(df) noise of sample_returns                  bootstrap.py:400 */
  int32_t* start_regime;         /* [n] 0 / 1, random_start.py:31 / :106 (NULL or zeros: regime 0) */
  double* boot_syn;              /* [max_blocks][2][boot_block_size][n] t(df) of the synthetic blocks that stand in for an
                                    empty pool (pool A / B of the block's regime), bootstrap.py:295; NULL when every
                                    pool has blocks */
} LetfTape;

/* per-trial outputs, structure-of-arrays: [strategy][n_trials] */
typedef struct LetfOutputs {
  int64_t n_trials;              /* capacity (columns) */
  double* final_wealth;          /* [S][n]  equity_curve[-1], engine.py:578 */
  double* max_dd;                /* [S][n]  engine.py:581-586 */
  int32_t* num_trades;           /* [S][n]  engine.py:589 */
  int32_t* journal_count;        /* [S][n]  len(Trade_List) for taxable ids (trade.py:63), may be NULL */
  double* journal_volume;        /* [S][n]  sum of dollar_amount (trade.py:75-81), may be NULL */
  int32_t* start_offset;         /* [n] */
  uint32_t* flags;               /* [n]  LETF_FLAG_* */
  int32_t* regime_days;          /* [2][n] Metadata['regime_counts'] over the window, may be NULL */
  int32_t* jump_days;            /* [n]  Metadata['stress_summary']['jump_days'], may be NULL */
} LetfOutputs;

/* summary buffer produced by letf_reduce: everything here is sum/min/max-reducible
 * across GPUs (NCCL allreduce over NVLink), SURVEY.md section 8(e). */
#define LETF_HIST_BINS 4096
#define LETF_N_METRICS 3        /* 0 log10(W/W0), 1 CAGR, 2 max drawdown */
typedef struct LetfSummary {
  /* all device pointers */
  unsigned long long* hist;      /* [S][3][LETF_HIST_BINS] */
  double* moments;               /* [S][3][2]  sum x, sum x^2 */
  double* extrema;               /* [S][3][2]  min, max (reduce with min / max) */
  unsigned long long* counters;  /* [S][4]  valid trials, wins vs strategy `bench`, total trades, flagged trials */
  double hist_lo[LETF_N_METRICS];
  double hist_hi[LETF_N_METRICS];
  int32_t bench_strategy;        /* index of the win-rate benchmark (S2), or -1 */
  int32_t pad;
} LetfSummary;

typedef struct LetfEngine LetfEngine;   /* opaque */

typedef struct LetfLaunchInfo {
  int32_t grid, block;
  int32_t regs_per_thread;
  int32_t kernel_launches;       /* path kernels launched by the last letf_run / letf_run_host / replay call */
  int64_t scratch_bytes;
  float last_kernel_ms;          /* CUDA-event time of the fused kernel in the last call */
  int32_t sm_count;
  int32_t aux_launches;          /* reduction / reset kernels launched by the last letf_reduce / letf_summary_reset */
  int32_t pad;
  char kernel[48];               /* instantiation name of the last path kernel (e.g. "f32_joint_t_pair_c3"; a split pipeline
                                    reports "<producer>+<rule kernel>", e.g. "joint_t_prod_sweep19+rules_sub19") */
} LetfLaunchInfo;

/* lifecycle ----------------------------------------------------------------
 * Tuning aids, read from the environment ONCE per engine by letf_create (they change which instantiation runs, never a result
 * beyond fp32 reassociation; the tests use them for A/B comparisons on the same Philox counters):
 *   LETF_KERNEL=<instantiation name>   pin an eligible build (LetfLaunchInfo::kernel reports what ran)
 *   LETF_NO_JPAIR=1                    buy&hold shapes keep the fused lean builds (also switches the routed / generic pipelines off)
 *   LETF_NO_SPLIT=1                    no split strategy pipeline at all (simulated paths and materialised-path mode)
 *   LETF_NO_ROUTED_SPLIT=1             subsets of the default sweep skip the routed 19-rule kernel (generic pipeline instead)
 *   LETF_NO_GENERIC_SPLIT=1            untyped strategy lists keep the fused generic builds
 *   LETF_TRACE_HOST=1                  host time of each step of the split launch on stderr
 *   LETF_SPLIT_WAVES=<1..4>, LETF_SERIES_CHUNK=<paths>   smaller chunks of the split launches (the tests cross chunk seams with them) */
int letf_create(const LetfParams* params, int device, int precision, LetfEngine** out);
void letf_destroy(LetfEngine* e);
const char* letf_last_error(void);
int letf_abi_version(void);

/* The fused path kernel.  `stream` is a cudaStream_t passed as void*.  All
 * pointers inside `tape` / `out` are DEVICE pointers.  Trials
 * [first_trial, first_trial + n_trials) are simulated and written to columns
 * [0, n_trials) of `out`.  rng_mode: LETF_RNG_*.  Replaces the loky fan-out of
 * mc_runner.py:155-165 plus everything under engine.py:423-619. */
int letf_run(LetfEngine* e, void* stream, uint64_t seed, int64_t first_trial,
             int64_t n_trials, int rng_mode, const LetfTape* tape, const LetfOutputs* out);

/* Same, with HOST buffers: allocates / reuses device staging, copies the tape
 * (if any) host->device, runs, copies results device->host, synchronises. */
int letf_run_host(LetfEngine* e, uint64_t seed, int64_t first_trial, int64_t n_trials,
                  int rng_mode, const LetfTape* host_tape, const LetfOutputs* host_out);

/* Histogram / moment reduction of per-trial outputs (device pointers).  letf_reduce ACCUMULATES
 * into `summary` (so several launches / trial ranges can be folded); letf_summary_reset zeroes the
 * bins, moments and counters and sets the extrema to +/-inf. */
int letf_summary_reset(LetfEngine* e, void* stream, const LetfSummary* summary);

/* The exchange step of a multi-GPU run (reporting.py:358-376 consumes the result).  A summary can live in ONE
 * packed device buffer of 8-byte words, [hist S*3*LETF_HIST_BINS | counters S*4 | moments S*3*2 | extrema S*3*2]
 * (letf_summary_words(S) words; the LetfSummary pointers then point into it), so that the ranks exchange it with a single
 * all-gather.  letf_summary_combine folds `world` gathered packed summaries (device, [world][words]) into `dst_packed`
 * (device, words): bins / counters / moments summed, extrema min / max.  One kernel, no host round trip. */
int64_t letf_summary_words(int n_strategies);
int letf_summary_combine(LetfEngine* e, void* stream, const void* gathered, int world, int n_strategies,
                         void* dst_packed);
int letf_reduce(LetfEngine* e, void* stream, const LetfOutputs* out, int64_t n_trials,
                const LetfSummary* summary);

/* Trade_List of ONE trial / ONE taxable strategy (trade.py:83-91), host buffers of capacity `cap`:
 * the trial is re-simulated (same seed or same tape) with the journal log switched on.  This is
 * what the host tax path needs for its <= 50 sampled trials per strategy (reporting.py:457-467).
 * action: +1 BUY, -1 SELL.  *count receives the number of entries (may exceed cap). */
int letf_trade_list_host(LetfEngine* e, uint64_t seed, int64_t trial, int strategy, int rng_mode,
                         const LetfTape* host_tape, int cap, int32_t* day, int32_t* asset,
                         int32_t* action, double* shares, double* price, int32_t* count);

/* Same for SEVERAL trials of one taxable strategy in ONE launch - the <= 50 wealth-ranked trials per strategy of
 * reporting.py:457-467: trial trials[j] is re-simulated by its own thread with the journal log on (Philox draws).
 * Host buffers day / asset / action / shares / price are [n][cap]; count[n] receives each trial's number of entries
 * (an entry count above cap means that trial's list was truncated: call again with a larger cap). */
int letf_trade_lists_host(LetfEngine* e, uint64_t seed, const int64_t* trials, int n, int strategy, int cap,
                          int32_t* day, int32_t* asset, int32_t* action, double* shares, double* price,
                          int32_t* count);

/* Same across SEVERAL strategies: item j re-simulates trial trials[j] and logs strategy strategies[j].  The whole sampled-trial
 * tax table (13 taxable strategies x 50 trials, reporting.py:408-560) is then ONE launch: the replay is bound by the latency of
 * a single thread walking its path, so 650 items cost what 50 do.  Items whose strategy is not taxable come back empty. */
int letf_trade_lists_multi_host(LetfEngine* e, uint64_t seed, const int64_t* trials, const int32_t* strategies, int n, int cap,
                                int32_t* day, int32_t* asset, int32_t* action, double* shares, double* price,
                                int32_t* count);

/* Daily equity curve of ONE trial / ONE strategy - the Series run_strategy_fixed returns (strategy.py:55-1277; its last
 * element is Final_Wealth, engine.py:578): the trial is re-simulated with the curve log on.  curve: host double[T]. */
int letf_equity_curve_host(LetfEngine* e, uint64_t seed, int64_t trial, int strategy, int rng_mode,
                           const LetfTape* host_tape, double* curve);

/* The generated underlying returns (SPY, QQQ, TLT; before the jump overlay) of ONE trial's simulated window - what
 * simulate_joint_returns_t(n_days, regime_path, joint_model, rng, antithetic) returns (calibration.py:480-590), moment
 * matching included: the trial is re-simulated (same seed or same tape) with the market log on.
 * returns: host double[sim_days][3].  With buffer_days == 0 the window is the whole generated path. */
int letf_market_returns_host(LetfEngine* e, uint64_t seed, int64_t trial, int rng_mode,
                             const LetfTape* host_tape, double* returns);

int letf_launch_info(LetfEngine* e, LetfLaunchInfo* info);

/* Strategy rules on caller-supplied return series (materialised-path mode):
 * the GPU counterpart of run_strategy_fixed(df, sid, regime_path, ...),
 * strategy.py:55.  series are [T][n] day-major device arrays. */
typedef struct LetfSeries {
  int64_t n_paths;
  int32_t sim_days;
  int32_t pad;
  const double* asset_ret[LETF_N_ASSETS]; /* <asset>_Ret, NULL if absent */
  const double* tlt_ret;
  const double* vix;
  const int32_t* regime;
  const double* cash_ret;                 /* Cash_Ret column (strategy.py:138 ...); NULL: the regime's cash rate / 252 (engine.py:544) */
} LetfSeries;
int letf_run_strategies(LetfEngine* e, void* stream, const LetfSeries* series,
                        const LetfOutputs* out);

/* run_strategy_fixed(df, sid, regime_path, corr, apply_costs, trade_journal) on ONE caller-supplied frame: `host_series`
 * holds HOST pointers to [T] columns (n_paths must be 1).  Returns the equity curve (host double[T]), the trade count and,
 * for a taxable id, the journal entries (buffers of capacity cap as in letf_trade_list_host; may be NULL with cap 0). */
int letf_run_strategy_curve_host(LetfEngine* e, const LetfSeries* host_series, int strategy, double* curve,
                                 int32_t* num_trades, int cap, int32_t* day, int32_t* asset, int32_t* action,
                                 double* shares, double* price, int32_t* count);


/* helper for non-Python callers: utils.py:42-53 on a 3x3 matrix */
int letf_nearest_psd3(const double in[9], double out[9]);

/* Host-only (no GPU work): realised gains / losses per tax year of ONE trade list - wash-sale basis adjustments, lot
 * selection and the short / long-term split - i.e. the per-trade bookkeeping of process_trades_with_wired_engine
 * (letf/integration.py:113-226 over letf/tax/wash_sale.py:149-397, letf/tax/lot_selection.py).  The <= 50 trade lists per
 * taxable strategy that create_summary_statistics taxes (letf/reporting.py:457-487) hold thousands of trades each at long
 * horizons; the yearly netting / bracket arithmetic on top of this table stays in Python (tax_engine.py).
 *   day / asset / action / shares / price: the journal in its own order, n entries (action 1 = BUY, 0 = SELL, others ignored;
 *   asset = small non-negative ids); lot_method: 0 fifo, 1 lifo, 2 hifo, 3 lofo, 4 ltfo, 5 stfo, 6 mintax, 7 specific
 *   (tax/engine.py:45-52); out: [n_years][5] = st_gains, st_losses, lt_gains, lt_losses, wash_sale_disallowed per year
 *   (day / 252). */
int letf_tax_realised_by_year(const int32_t* day, const int32_t* asset, const int32_t* action, const double* shares,
                              const double* price, int64_t n, int lot_method, int n_years, double* out);

#ifdef __cplusplus
}
#endif
#endif /* LETF_B200_H */
