// Host emulation of the fused kernel's per-trial body (TEST-ONLY build).
//
// letf_core.cuh is written __host__ __device__; compiling it with the host compiler lets the
// CPU test-suite exercise the *kernel source* (fp64 and fp32 instantiations, tape / Philox /
// record modes) against the oracle without a GPU.  It is never part of the product: the
// shipped library contains only the device kernels and has no CPU path.
#include <string.h>
#include <stdlib.h>
#include <vector>
#include "../../letf_monte_carlo_b200/csrc/letf_core.cuh"
#include "../../letf_monte_carlo_b200/csrc/letf_fast.cuh"
#include "../../letf_monte_carlo_b200/csrc/letf_jfast.cuh"
#include "../../letf_monte_carlo_b200/csrc/letf_series_prep.cuh"

using namespace letf;

template <typename R, int MARKET>
static int run_t(const LetfParams* sp, const LetfTape* tape, const LetfOutputs* out, uint64_t seed,
                 int64_t first, int64_t n, int rng_mode, int64_t log_trial, int log_strategy, int log_cap,
                 int32_t* log_count, int32_t* log_day, int32_t* log_asset, int32_t* log_action,
                 double* log_shares, double* log_price) {
  constexpr int NS = LETF_MAX_STRATEGIES;
  constexpr int NA = kAssets;
  static KArgs<R, NS, NA> ka;
  const int32_t* dd[2] = {sp->duration_samples[0], sp->duration_samples[1]};
  static BootTables<R> boot;                       // host copies of the bootstrap tables (device blob in the product)
  build_boot_tables<R>(*sp, boot);
  if (sp->market_mode == LETF_MARKET_BOOTSTRAP) {
    for (int c = 0; c < 3; c++) boot.meta.hist[c] = boot.hist.data() + (size_t)c * sp->boot.hist_len;
    boot.meta.cstart = boot.cstart.data(); boot.meta.cret = boot.cret.data();
  }
  convert_params<R, NS, NA>(*sp, dd, ka.p, &boot.meta);
  if (tape) ka.tape = *tape; else memset(&ka.tape, 0, sizeof(ka.tape));
  ka.out = *out;
  ka.seed = seed; ka.first_trial = first; ka.n_trials = n;
  ka.tape_first = 0;
  ka.rng_mode = rng_mode;
  const int G = sp->sim_days + sp->buffer_days;
  std::vector<R> xscr((size_t)(G > 0 ? G : 1) * 5 + 8), rp(kRingP), rs(kRingS), ra(kRingA);
  ka.xscr = xscr.data(); ka.ringP = rp.data(); ka.ringS = rs.data(); ka.ringA = ra.data();
  ka.nthreads = 1;
  ka.log_trial = log_trial; ka.log_strategy = log_strategy; ka.log_cap = log_cap;
  ka.log_count = log_count; ka.log_day = log_day; ka.log_asset = log_asset; ka.log_action = log_action;
  ka.log_shares = log_shares; ka.log_price = log_price;
  for (int64_t j = 0; j < n; j++) {
    Trial<R, NS, NA, MARKET, true, 0xFFFFFFFFu> t(ka, first + j, j, 0);
    t.run();
  }
  return 0;
}

// the simulate_joint_returns_t seam: generated SPY / QQQ / TLT returns of one trial's window (kernel source, market log on)
template <typename R, int MARKET>
static int market_t(const LetfParams* sp, const LetfTape* tape, uint64_t seed, int64_t trial, int rng_mode, double* returns) {
  constexpr int NS = LETF_MAX_STRATEGIES;
  constexpr int NA = kAssets;
  static KArgs<R, NS, NA> ka;
  memset(&ka, 0, sizeof(ka));
  const int32_t* dd[2] = {sp->duration_samples[0], sp->duration_samples[1]};
  convert_params<R, NS, NA>(*sp, dd, ka.p);
  if (tape) ka.tape = *tape;
  const int S = sp->n_strategies > 0 ? sp->n_strategies : 1;
  std::vector<double> fw(S), md(S); std::vector<int32_t> nt(S);
  ka.out.n_trials = 1; ka.out.final_wealth = fw.data(); ka.out.max_dd = md.data(); ka.out.num_trades = nt.data();
  ka.seed = seed; ka.first_trial = trial; ka.n_trials = 1; ka.rng_mode = rng_mode;
  const int G = sp->sim_days + sp->buffer_days;
  std::vector<R> xscr((size_t)(G > 0 ? G : 1) * 5 + 8), rp(kRingP), rs(kRingS), ra(kRingA);
  ka.xscr = xscr.data(); ka.ringP = rp.data(); ka.ringS = rs.data(); ka.ringA = ra.data();
  ka.nthreads = 1; ka.log_trial = trial; ka.log_strategy = -1; ka.log_market = returns;
  Trial<R, NS, NA, MARKET, true, 0xFFFFFFFFu> t(ka, trial, 0, 0);
  t.run();
  return 0;
}
extern "C" int letf_emu_market(const LetfParams* sp, int precision, const LetfTape* tape, uint64_t seed, int64_t trial, int rng_mode,
                               double* returns) {
  if (validate_params(*sp) || sp->market_mode != LETF_MARKET_JOINT_T) return -1;
  if (precision == LETF_PREC_F64) return market_t<double, 2>(sp, tape, seed, trial, rng_mode, returns);
  return market_t<float, 2>(sp, tape, seed, trial, rng_mode, returns);
}

extern "C" int letf_emu_run(const LetfParams* sp, int precision, const LetfTape* tape, const LetfOutputs* out,
                            uint64_t seed, int64_t first, int64_t n, int rng_mode,
                            int64_t log_trial, int log_strategy, int log_cap, int32_t* log_count,
                            int32_t* log_day, int32_t* log_asset, int32_t* log_action, double* log_shares,
                            double* log_price) {
  const char* err = validate_params(*sp);
  if (err) return -1;
#define GO(R, MK) return run_t<R, MK>(sp, tape, out, seed, first, n, rng_mode, log_trial, log_strategy, log_cap, log_count, log_day, log_asset, log_action, log_shares, log_price)
  if (precision == LETF_PREC_F64) {
    if (sp->market_mode == 0) GO(double, 0); else if (sp->market_mode == 1) GO(double, 1);
    else if (sp->market_mode == LETF_MARKET_BOOTSTRAP) GO(double, 4); else GO(double, 2);
  } else {
    if (sp->market_mode == 0) GO(float, 0); else if (sp->market_mode == 1) GO(float, 1);
    else if (sp->market_mode == LETF_MARKET_BOOTSTRAP) GO(float, 4); else GO(float, 2);
  }
#undef GO
}

// the lean buy&hold build (compacted asset slots, loop-unswitched FAST day step), Philox only
extern "C" int letf_emu_run_lean(const LetfParams* sp, const LetfOutputs* out, uint64_t seed, int64_t first, int64_t n,
                                 int* used_fast, int two_per_thread, int force_generic) {
  if (validate_params(*sp) || sp->market_mode != LETF_MARKET_NORMAL) return -1;
  constexpr int NS = 4, NA = 3;
  constexpr uint32_t TMB = (1u << LETF_ST_BENCHMARK) | (1u << LETF_ST_BENCHMARK_TLT);
  if (sp->n_strategies > NS || __builtin_popcount((unsigned)sp->asset_mask & 0x3Fu) > NA) return -2;
  static KArgs<float, NS, NA> ka;
  memset(&ka, 0, sizeof(ka));
  const int32_t* dd[2] = {nullptr, nullptr};
  convert_params<float, NS, NA>(*sp, dd, ka.p);
  ka.out = *out; ka.seed = seed; ka.first_trial = first; ka.n_trials = n; ka.rng_mode = LETF_RNG_PHILOX; ka.nthreads = 1;
  ka.log_trial = -1;
  const bool fast_ok = fast_pattern_compiled<NA>(ka.p.fast_layers);
  if (used_fast) *used_fast = fast_ok ? ka.p.fast_layers : 0;
  if (two_per_thread && fast_ok && !force_generic) {
    // the launch form of the fast path: two interleaved trials per thread (letf_fast.cuh), incl. an odd tail
    for (int64_t j = 0; j < n; j += 2) {
      int64_t trial[2] = {first + j, first + (j + 1 < n ? j + 1 : j)}, col[2] = {j, j + 1 < n ? j + 1 : j};
      bool live[2] = {true, j + 1 < n};
      fast_layers_run<NS, NA, 2>(ka, trial, col, live);
    }
    return 0;
  }
  if (fast_ok && !two_per_thread && !force_generic) {
    for (int64_t j = 0; j < n; j++) {
      int64_t trial[1] = {first + j}, col[1] = {j};
      bool live[1] = {true};
      fast_layers_run<NS, NA, 1>(ka, trial, col, live);
    }
    return 0;
  }
  for (int64_t j = 0; j < n; j++) {
    Trial<float, NS, NA, LETF_MARKET_NORMAL, false, TMB> t(ka, first + j, j, 0);
    t.run();
  }
  return 0;
}

// the lean institutional buy&hold build: the antithetic-pair kernel body (letf_jfast.cuh) when the shape is admitted and
// compiled in, else (or with force_generic) the lean build's general loop; Philox only
template <int NACT, int K10>
static void run_jpairs(const KArgs<float, 4, 3>& ka, int64_t first, int64_t n) {
  for (int64_t pj = first >> 1; pj <= (first + n - 1) >> 1; pj++) JointPair<4, 3, NACT, K10>::run(ka, pj, 0);
}
extern "C" int letf_emu_run_jpair(const LetfParams* sp, const LetfOutputs* out, uint64_t seed, int64_t first, int64_t n,
                                  int* used_pair, int force_generic) {
  if (validate_params(*sp) || sp->market_mode != LETF_MARKET_JOINT_T) return -1;
  constexpr int NS = 4, NA = 3;
  constexpr uint32_t TMB = (1u << LETF_ST_BENCHMARK) | (1u << LETF_ST_BENCHMARK_TLT);
  if (sp->n_strategies > NS || __builtin_popcount((unsigned)sp->asset_mask & 0x3Fu) > NA) return -2;
  static KArgs<float, NS, NA> ka;
  memset(&ka, 0, sizeof(ka));
  const int32_t* dd[2] = {sp->duration_samples[0], sp->duration_samples[1]};
  convert_params<float, NS, NA>(*sp, dd, ka.p);
  ka.out = *out; ka.seed = seed; ka.first_trial = first; ka.n_trials = n; ka.rng_mode = LETF_RNG_PHILOX; ka.nthreads = 1;
  ka.log_trial = -1;
  const int G = sp->sim_days + sp->buffer_days;
  float* xscr = (float*)aligned_alloc(32, ((size_t)(G > 0 ? G : 1) * 8 + 8) * sizeof(float));
  ka.xscr = xscr;
  std::vector<uint32_t> spells((size_t)G + 2);
  ka.spell_scr = spells.data();
  const int code = ka.p.fast_joint;
  const bool compiled = code == fast_joint_code(2, 3) || code == fast_joint_code(3, 4) || code == fast_joint_code(1, 0) ||
                        code == fast_joint_code(1, 1) || code == fast_joint_code(2, 0x80) || code == fast_joint_code(3, 0x80);
  if (used_pair) *used_pair = (compiled && !force_generic) ? code : 0;
  if (compiled && !force_generic && n > 0) {
    if (code == fast_joint_code(2, 3)) run_jpairs<2, 3>(ka, first, n);
    else if (code == fast_joint_code(3, 4)) run_jpairs<3, 4>(ka, first, n);
    else if (code == fast_joint_code(1, 0)) run_jpairs<1, 0>(ka, first, n);
    else if (code == fast_joint_code(2, 0x80)) run_jpairs<2, 0x80>(ka, first, n);
    else if (code == fast_joint_code(3, 0x80)) run_jpairs<3, 0x80>(ka, first, n);
    else run_jpairs<1, 1>(ka, first, n);
  } else {
    for (int64_t j = 0; j < n; j++) {
      Trial<float, NS, NA, LETF_MARKET_JOINT_T, false, TMB> t(ka, first + j, j, 0);
      t.run();
    }
  }
  free(xscr);
  return 0;
}

// typed-slot builds (one compile-time rule per slot, compile-time asset mask), Philox only, fp32
template <int NS, int LIST, int AM>
static int run_typed(const LetfParams* sp, const LetfOutputs* out, uint64_t seed, int64_t first, int64_t n) {
  constexpr int NA = kAssets;
  if (sp->n_strategies != slot_list_count(LIST) || sp->asset_mask != AM || sp->market_mode != LETF_MARKET_JOINT_T) return -2;
  for (int s = 0; s < sp->n_strategies; s++) if (sp->strategies[s].type != slot_list_type(LIST, s)) return -3;
  static KArgs<float, NS, NA> ka;
  memset(&ka, 0, sizeof(ka));
  const int32_t* dd[2] = {sp->duration_samples[0], sp->duration_samples[1]};
  convert_params<float, NS, NA>(*sp, dd, ka.p);
  ka.out = *out; ka.seed = seed; ka.first_trial = first; ka.n_trials = n; ka.rng_mode = LETF_RNG_PHILOX; ka.nthreads = 1;
  ka.log_trial = -1;
  const int G = sp->sim_days + sp->buffer_days;
  std::vector<float> xscr((size_t)(G > 0 ? G : 1) * 5 + 8), rp(kRingP), rs(kRingS), ra(kRingA);
  ka.xscr = xscr.data(); ka.ringP = rp.data(); ka.ringS = rs.data(); ka.ringA = ra.data();
  for (int64_t j = 0; j < n; j++) {
    Trial<float, NS, NA, LETF_MARKET_JOINT_T, false, slot_list_mask(LIST), LIST, AM> t(ka, first + j, j, 0);
    t.run();
  }
  return 0;
}
extern "C" int letf_emu_run_typed(const LetfParams* sp, int list, const LetfOutputs* out, uint64_t seed, int64_t first,
                                  int64_t n) {
  if (validate_params(*sp)) return -1;
  if (list == 1) return run_typed<8, 1, 25>(sp, out, seed, first, n);
  if (list == 2) return run_typed<32, 2, 29>(sp, out, seed, first, n);
  return -4;
}

// split strategy pipeline (letf_api.cu launch_split): the pair producer writes per-day records, the typed rule kernel body
// (Trial::run in record mode) consumes them; Philox only, fp32
template <int NS, int LIST, int AM, int NACT, int K10, int NAP = 4>
static int run_split(const LetfParams* sp, const LetfOutputs* out, uint64_t seed, int64_t first, int64_t n) {
  if (sp->market_mode != LETF_MARKET_JOINT_T) return -2;
  // routed list (letf_api.cu find_routed_split): the request is a subset of the typed list
  static LetfParams routed;
  int slot_of[LETF_MAX_STRATEGIES], n_slots = 0;
  const LetfParams* req = sp;
  if (slot_list_routed(LIST)) {
    bool rules = false;
    if ((sp->asset_mask & 0x3F) & ~AM) return -2;
    if (!route_request(*sp, LIST, NS, slot_of, &rules)) return -3;
    n_slots = routed_slot_count(*sp, LIST, slot_of);
    routed = routed_params(*sp, LIST, slot_of, n_slots, AM);
    sp = &routed;
  } else if (LIST != 0) {
    if (sp->n_strategies != slot_list_count(LIST) || sp->asset_mask != AM) return -2;
    for (int s = 0; s < sp->n_strategies; s++) if (sp->strategies[s].type != slot_list_type(LIST, s)) return -3;
  }
  constexpr uint32_t TMC = LIST != 0 ? slot_list_mask(LIST) : ((1u << LETF_ST__COUNT) - 1u);   // letf_kernels.cuh TM_ALL
  static KArgs<float, 32, NAP> kp;
  static KArgs<float, NS, kAssets> kc;
  memset(&kp, 0, sizeof(kp)); memset(&kc, 0, sizeof(kc));
  const int32_t* dd[2] = {sp->duration_samples[0], sp->duration_samples[1]};
  convert_params<float, 32, NAP>(*sp, dd, kp.p);
  convert_params<float, NS, kAssets>(*sp, dd, kc.p);
  if (slot_list_routed(LIST)) { route_slots(kp.p, *req, slot_of, n_slots); route_slots(kc.p, *req, slot_of, n_slots); }
  const int T = sp->sim_days, G = T + sp->buffer_days, NF = NACT + 2;
  const int64_t cap = ((n + 1) & ~(int64_t)1) + 2;        // pair-aligned columns + the half pairs at both ends (letf_api.cu launch_split)
  const int odd = (int)(first & 1);
  std::vector<float> rec((size_t)cap * T * NF + 8), rp(kRingP), rs(kRingS), ra(kRingA);
  float* xscr = (float*)aligned_alloc(32, ((size_t)(G > 0 ? G : 1) * 8 + 8) * sizeof(float));
  kp.out = *out; kp.seed = seed; kp.first_trial = first; kp.n_trials = n; kp.rng_mode = LETF_RNG_PHILOX; kp.nthreads = 1;
  std::vector<uint32_t> spells((size_t)G + 2);
  kp.xscr = xscr; kp.rec = rec.data(); kp.rec_ncol = cap; kp.log_trial = -1; kp.spell_scr = spells.data();
  for (int64_t pj = first >> 1; pj <= (first + n - 1) >> 1; pj++) JointPair<32, NAP, NACT, K10, true>::run(kp, pj, 0);
  kc.out = *out; kc.seed = seed; kc.first_trial = first; kc.n_trials = n; kc.rng_mode = LETF_RNG_PHILOX; kc.nthreads = 1;
  kc.rec = rec.data() + odd; kc.rec_ncol = cap; kc.log_trial = -1;
  kc.p.T = T; kc.p.buffer = 0; kc.p.random_start = 0; kc.series.sim_days = T; kc.series.n_paths = n;
  kc.ringP = rp.data(); kc.ringS = rs.data(); kc.ringA = ra.data();
  for (int64_t j = 0; j < n; j++) {
    Trial<float, NS, kAssets, kMarketFSeries, false, TMC, LIST, AM> t(kc, first + j, j, 0);
    t.run();
  }
  free(xscr);
  return 0;
}
extern "C" int letf_emu_run_split(const LetfParams* sp, int list, const LetfOutputs* out, uint64_t seed, int64_t first, int64_t n) {
  if (validate_params(*sp)) return -1;
  bool tedf = false;
  for (int a = 0; a < kAssets; a++)
    if (((sp->asset_mask >> a) & 1) && ((df_fast_k(sp->assets[a].te_df) != 5 && df_fast_k(sp->assets[a].te_df) != 10) || sp->assets[a].te_clip > 0.0)) tedf = true;
  if (list == 1) return tedf ? run_split<8, 1, 25, 3, 0x80>(sp, out, seed, first, n) : run_split<8, 1, 25, 3, 4>(sp, out, seed, first, n);
  if (list == 2) return tedf ? run_split<32, 2, 29, 4, 0x80>(sp, out, seed, first, n) : run_split<32, 2, 29, 4, 8>(sp, out, seed, first, n);
  if (list == 4) {      // routed: a subset of the config-4 list on its six-slot kernel
    tedf = false;
    for (int a = 0; a < kAssets; a++)
      if (((25 >> a) & 1) && ((df_fast_k(sp->assets[a].te_df) != 5 && df_fast_k(sp->assets[a].te_df) != 10) || sp->assets[a].te_clip > 0.0)) tedf = true;
    return tedf ? run_split<8, 4, 25, 3, 0x80>(sp, out, seed, first, n) : run_split<8, 4, 25, 3, 4>(sp, out, seed, first, n);
  }
  if (list == 3) {      // routed: a subset of the default sweep on the 19-rule kernel with the other slots switched off
    tedf = false;
    for (int a = 0; a < kAssets; a++)
      if (((29 >> a) & 1) && ((df_fast_k(sp->assets[a].te_df) != 5 && df_fast_k(sp->assets[a].te_df) != 10) || sp->assets[a].te_clip > 0.0)) tedf = true;
    return tedf ? run_split<32, 3, 29, 4, 0x80>(sp, out, seed, first, n) : run_split<32, 3, 29, 4, 8>(sp, out, seed, first, n);
  }
  if (list == 0) {      // generic pipeline: uncompacted producer (default dof pattern or run-time sampler), dynamic-slot rule kernel
    for (int a = 0; a < kAssets; a++)
      if (((sp->asset_mask >> a) & 1) && df_fast_k(sp->assets[a].te_df) != (a >= 4 ? 10 : 5)) tedf = true;
    if (sp->n_strategies <= 8)      // (the eight-slot build: slot loop unrolled, state in registers)
      return tedf ? run_split<8, 0, 0x3F, 6, 0x80, 6>(sp, out, seed, first, n) : run_split<8, 0, 0x3F, 6, 0x30, 6>(sp, out, seed, first, n);
    return tedf ? run_split<32, 0, 0x3F, 6, 0x80, 6>(sp, out, seed, first, n) : run_split<32, 0, 0x3F, 6, 0x30, 6>(sp, out, seed, first, n);
  }
  return -4;
}

// materialised-path mode (run_strategy_fixed seam): caller-supplied series, optional journal + equity-curve log of one
// (path, strategy), optional batch replay ids.  typed = 1 runs the typed 19-rule fp32 series build.
template <typename R, int NS, bool TAPEIO, uint32_t TM, int LIST>
static int run_series_t(const LetfParams* sp, const LetfSeries* series, const LetfOutputs* out, int64_t log_trial, int log_strategy,
                        int log_cap, int32_t* log_count, int32_t* log_day, int32_t* log_asset, int32_t* log_action,
                        double* log_shares, double* log_price, double* log_curve) {
  constexpr int NA = kAssets;
  static KArgs<R, NS, NA> ka;
  memset(&ka, 0, sizeof(ka));
  const int32_t* dd[2] = {nullptr, nullptr};
  convert_params<R, NS, NA>(*sp, dd, ka.p);
  ka.series = *series; ka.p.T = series->sim_days; ka.p.buffer = 0; ka.p.random_start = 0;
  ka.out = *out; ka.n_trials = series->n_paths; ka.rng_mode = LETF_RNG_PHILOX; ka.nthreads = 1;
  std::vector<R> rp(kRingP), rs(kRingS), ra(kRingA);
  ka.ringP = rp.data(); ka.ringS = rs.data(); ka.ringA = ra.data();
  ka.log_trial = log_trial; ka.log_strategy = log_strategy; ka.log_cap = log_cap;
  ka.log_count = log_count; ka.log_day = log_day; ka.log_asset = log_asset; ka.log_action = log_action;
  ka.log_shares = log_shares; ka.log_price = log_price; ka.log_curve = log_curve;
  for (int64_t j = 0; j < series->n_paths; j++) {
    Trial<R, NS, NA, 3, TAPEIO, TM, LIST, 0> t(ka, j, j, 0);
    t.run();
  }
  return 0;
}
extern "C" int letf_emu_run_series(const LetfParams* sp, int precision, int typed, const LetfSeries* series, const LetfOutputs* out,
                                   int64_t log_trial, int log_strategy, int log_cap, int32_t* log_count, int32_t* log_day,
                                   int32_t* log_asset, int32_t* log_action, double* log_shares, double* log_price, double* log_curve) {
  if (validate_params(*sp)) return -1;
#define GO(R, NS, IO, TMASK, LST) return run_series_t<R, NS, IO, TMASK, LST>(sp, series, out, log_trial, log_strategy, log_cap, log_count, log_day, log_asset, log_action, log_shares, log_price, log_curve)
  if (typed) {
    if (precision != LETF_PREC_F32 || sp->n_strategies != slot_list_count(2)) return -2;
    GO(float, 32, false, slot_list_mask(2), 2);
  }
  if (precision == LETF_PREC_F64) GO(double, LETF_MAX_STRATEGIES, true, 0xFFFFFFFFu, 0);
  GO(float, LETF_MAX_STRATEGIES, true, 0xFFFFFFFFu, 0);
#undef GO
}

// materialised-path mode through the split pipeline (letf_api.cu launch_series_split): record pass over the series, then the
// rule kernel body in record mode.  list: 1 / 2 typed, 3 routed, 0 dynamic slots
template <int NS, int LIST, int AM>
static int run_series_split(const LetfParams* sp, const LetfSeries* series, const LetfOutputs* out) {
  static LetfParams routed;
  int slot_of[LETF_MAX_STRATEGIES], n_slots = 0;
  const LetfParams* req = sp;
  if (slot_list_routed(LIST)) {
    bool rules = false;
    if ((sp->asset_mask & 0x3F) & ~AM) return -2;
    if (!route_request(*sp, LIST, NS, slot_of, &rules)) return -3;
    n_slots = routed_slot_count(*sp, LIST, slot_of);
    routed = routed_params(*sp, LIST, slot_of, n_slots, AM);
    sp = &routed;
  } else if (LIST != 0) {
    if (sp->n_strategies != slot_list_count(LIST) || sp->asset_mask != AM) return -2;
    for (int s = 0; s < sp->n_strategies; s++) if (sp->strategies[s].type != slot_list_type(LIST, s)) return -3;
  }
  constexpr uint32_t TMC = LIST != 0 ? slot_list_mask(LIST) : ((1u << LETF_ST__COUNT) - 1u);
  static KArgs<float, NS, kAssets> kc;
  static SeriesPrepArgs pa;
  memset(&kc, 0, sizeof(kc)); memset(&pa, 0, sizeof(pa));
  const int32_t* dd[2] = {nullptr, nullptr};
  convert_params<float, NS, kAssets>(*sp, dd, kc.p);
  if (slot_list_routed(LIST)) route_slots(kc.p, *req, slot_of, n_slots);
  const int T = series->sim_days, NF = __builtin_popcount((unsigned)AM & 0x3Fu) + 2;
  const int64_t n = series->n_paths;
  std::vector<float> rec((size_t)n * T * NF + 8), rp(kRingP), rs(kRingS), ra(kRingA);
  pa.series = *series; pa.out = *out; pa.rec = rec.data(); pa.rec_ncol = n; pa.col0 = 0; pa.n = n; pa.T = T;
  std::vector<float> cash(series->cash_ret ? (size_t)n * T : 0);
  if (series->cash_ret) { pa.rec_cash = cash.data(); kc.rec_cash = cash.data(); }
  pa.field_mask = AM & 0x3F; pa.asset_mask = kc.p.asset_mask; pa.bench_mask = kc.p.bench_mask; pa.bench_tlt = kc.p.bench_tlt;
  pa.n_strat = kc.p.n_strat; pa.W0 = (double)kc.p.W0;
  for (int q = 0; q < kc.p.n_strat; q++) { pa.type[q] = kc.p.strat[q].type; pa.asset[q] = kc.p.strat[q].asset; pa.route[q] = kc.p.strat[q].route; }
  for (int64_t j = 0; j < n; j++) series_prep_path(pa, j);
  kc.out = *out; kc.n_trials = n; kc.rng_mode = LETF_RNG_PHILOX; kc.nthreads = 1; kc.rec = rec.data(); kc.rec_ncol = n; kc.log_trial = -1;
  kc.p.T = T; kc.p.buffer = 0; kc.p.random_start = 0; kc.series.sim_days = T; kc.series.n_paths = n;
  kc.ringP = rp.data(); kc.ringS = rs.data(); kc.ringA = ra.data();
  for (int64_t j = 0; j < n; j++) {
    Trial<float, NS, kAssets, kMarketFSeries, false, TMC, LIST, AM> t(kc, j, j, 0);
    t.run();
  }
  return 0;
}
extern "C" int letf_emu_run_series_split(const LetfParams* sp, int list, const LetfSeries* series, const LetfOutputs* out) {
  if (validate_params(*sp)) return -1;
  if (series->sim_days < 1) return -5;
  if ((list == 1 || list == 2) && series->cash_ret) list = 3;      // the typed rule kernels read no cash column (letf_api.cu find_series_split)
  if (list == 1) return run_series_split<8, 1, 25>(sp, series, out);
  if (list == 2) return run_series_split<32, 2, 29>(sp, series, out);
  if (list == 3) return run_series_split<32, 3, 29>(sp, series, out);
  if (list == 4) return run_series_split<8, 4, 25>(sp, series, out);
  if (list == 0) return sp->n_strategies <= 8 ? run_series_split<8, 0, 0x3F>(sp, series, out) : run_series_split<32, 0, 0x3F>(sp, series, out);
  return -4;
}

extern "C" const char* letf_emu_validate(const LetfParams* sp) { return validate_params(*sp); }
