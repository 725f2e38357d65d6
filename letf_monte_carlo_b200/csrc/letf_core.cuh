// letf_core.cuh — one Monte Carlo trial, streamed day by day with all state in registers
// (plus three per-thread history rings and, for moment matching, a per-thread return scratch
// in HBM).  This is the body of the fused kernel; reference call stack: SURVEY.md 3.2.
//
//   regime path        engine.py:255-302          RegimeGen
//   joint-t GARCH/DCC  calibration.py:480-590     JointState::step
//   parametric-t       engine.py:81-97
//   VIX / IRX          engine.py:100-123
//   stress + jumps     calibration.py:787-830, engine.py:502-512
//   financing          engine.py:320-349, calibration.py:833-857, utils.py:169-216
//   tracking error     engine.py:157-203
//   ETF composition    engine.py:360-420
//   frame + validity   engine.py:537-562
//   strategies         strategy.py:55-1277        StrategyBank
//   per-trial stats    engine.py:578-590
#pragma once
#include <type_traits>
#include <utility>
#include "letf_params.cuh"
#include "letf_rolling.cuh"

namespace letf {

template <typename R, int NS, int NA> struct KArgs {
  DevParams<R, NS, NA> p;
  LetfTape tape;            // device pointers (tape / record modes)
  LetfOutputs out;          // device pointers
  LetfSeries series;        // MARKET == 3 only: caller-supplied frame (run_strategy_fixed seam)
  uint64_t seed;
  int64_t first_trial, n_trials, tape_first;
  int rng_mode;
  int pad;
  R* xscr;                  // [nthreads][G][4]   two-pass scratch (joint-t moment matching, block bootstrap): one aligned
                            // 4-vector per generated day, thread-major (each lane streams its own column; window days
                            // are offset by the lane's random start, so a day-major layout cannot coalesce the second
                            // pass); behind it [nthreads][G rounded up to 4] scalars (Trial::run, AUXSCR)
  R* ringP;                 // [kRingP][nthreads]
  R* ringS;                 // [kRingS][nthreads]
  R* ringA;                 // [kRingA][nthreads]
  int64_t nthreads;
  // optional journal log of ONE strategy (Trade_List replay): of the single trial log_trial, or - when trial_ids is
  // set - of every launched trial (work item j simulates trial_ids[j] and logs into row j of the [n][log_cap] buffers)
  const int64_t* trial_ids;
  const int32_t* log_strategies;   // batch replay over several strategies: work item j logs strategy log_strategies[j] (nullptr: log_strategy)
  int64_t log_trial; int log_strategy; int log_cap;
  int32_t* log_count; int32_t* log_day; int32_t* log_asset; int32_t* log_action; double* log_shares; double* log_price;
  double* log_curve;        // [T] daily equity of (log_trial, log_strategy): the Series run_strategy_fixed returns
  double* log_market;       // [T][3] SPY / QQQ / TLT returns of log_trial's window (the simulate_joint_returns_t seam)
  // split strategy pipeline (letf_jfast.cuh EMIT producer -> Trial::run with MARKET == kMarketFSeries): per-day records,
  // rec[(field * T + day) * rec_ncol + column], fields = the active layer returns, VIX (regime in the sign bit), SPY price
  float* rec;
  int64_t rec_ncol;
  const float* rec_cash;    // record mode on a caller-supplied frame with its own Cash_Ret column: [day][rec_ncol] (letf_series_prep.cuh);
                            // nullptr = the regime's cash rate (engine.py:544), which is what the simulated frames carry
  uint32_t* spell_scr;      // [nthreads][G + 1]: regime spells of the pair kernels (letf_jfast.cuh), (end day << 1) | regime
};

LETF_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ------------------------------------------------------------------------------------------
// Typed strategy-slot lists: builds in which slot s runs ONE compile-time rule, so each rule's code
// exists once and no per-day type dispatch is left.  list 0 = dynamic dispatch on p.strat[s].type.
//   1: BASELINE config 4 sweep        S1 S2 S4 S5 S6 S19
//   2: the reference's default sweep  S1 ... S19 in cfg.STRATEGIES order (letf/config.py:348-437)
//   3: list 2 with routed slots: any subset of the default sweep - a slot can be switched off and writes the output row
//      DevStrategy::route names (record mode only; letf_api.cu find_routed_split)
//   4: list 1 with routed slots (subsets of the config-4 list: the cheaper six-slot kernel)
// -2 marks an empty slot.
LETF_HD constexpr bool slot_list_routed(int list) { return list == 3 || list == 4; }
LETF_HD constexpr int slot_list_type(int list, int s) {
  if (list == 1 || list == 4) {
    switch (s) {
      case 0: case 1: return LETF_ST_BENCHMARK;
      case 2: return LETF_ST_SMA;
      case 3: return LETF_ST_SMA_BAND;
      case 4: return LETF_ST_PORTFOLIO;
      case 5: return LETF_ST_CONVICTION;
      default: return -2;
    }
  }
  if (list == 2 || list == 3) {
    switch (s) {
      case 0: case 1: case 2: return LETF_ST_BENCHMARK;
      case 3: return LETF_ST_SMA;
      case 4: return LETF_ST_SMA_BAND;
      case 5: return LETF_ST_PORTFOLIO;
      case 6: return LETF_ST_VOL_TARGETING;
      case 7: return LETF_ST_COMPOSITE;
      case 8: return LETF_ST_ADAPTIVE_VOL;
      case 9: return LETF_ST_DOWNSIDE_VOL;
      case 10: return LETF_ST_CONVEX_VOL;
      case 11: return LETF_ST_VOL_VELOCITY;
      case 12: return LETF_ST_VOL_MOM;
      case 13: return LETF_ST_SKEW_CONVEX;
      case 14: return LETF_ST_META_ENSEMBLE;
      case 15: return LETF_ST_REGIME_ASYMMETRIC;
      case 16: return LETF_ST_SKEW_KELLY;
      case 17: return LETF_ST_MOM_VOL_CONV;
      case 18: return LETF_ST_CONVICTION;
      default: return -2;
    }
  }
  return -1;
}
// every rule of the typed lists trades TQQQ (S8's defensive leg: SPY), cfg.STRATEGIES (config.py:348-437);
// buy & hold slots read their asset at run time (they do no per-day work)
LETF_HD constexpr int slot_list_asset(int list, int s) {
  const int t = slot_list_type(list, s);
  return (list != 0 && t >= 0 && t != LETF_ST_BENCHMARK && t != LETF_ST_BENCHMARK_TLT && t != LETF_ST_PORTFOLIO) ? 0 : -1;
}
LETF_HD constexpr int slot_list_count(int list) {
  int n = 0;
  for (int s = 0; s < LETF_MAX_STRATEGIES; s++) if (slot_list_type(list, s) >= 0) n = s + 1;
  return n;
}
LETF_HD constexpr uint32_t slot_list_mask(int list) {
  uint32_t m = 0;
  for (int s = 0; s < LETF_MAX_STRATEGIES; s++) if (slot_list_type(list, s) >= 0) m |= 1u << slot_list_type(list, s);
  return m;
}
template <int LIST, typename F, int... S>
LETF_HD void for_typed_slots(F& f, std::integer_sequence<int, S...>) {
  (f(S, std::integral_constant<int, slot_list_type(LIST, S)>{}, std::integral_constant<int, slot_list_asset(LIST, S)>{}), ...);
}

// ---- routed typed lists (host side): any subset of a typed list runs that list's rule kernel with the other slots switched off.
// Every rule of the request goes to the slot that compiles its type, buy&hold strategies to the list's buy&hold slots or to
// the free slots behind it (they cost the rule kernel nothing: the producer writes them).  slot_of[q] = slot of requested
// strategy q; false = some strategy has no slot (a rule type twice, a rule on another asset, more strategies than slots).
inline bool route_request(const LetfParams& s, int list, int ns, int* slot_of, bool* has_rules) {
  const int nl = slot_list_count(list);
  bool used[LETF_MAX_STRATEGIES] = {false};
  *has_rules = false;
  for (int q = 0; q < s.n_strategies; q++) {
    const LetfStrategyParams& st = s.strategies[q];
    const bool hold = st.type == LETF_ST_BENCHMARK || st.type == LETF_ST_BENCHMARK_TLT;
    int slot = -1;
    for (int j = 0; j < ns && j < LETF_MAX_STRATEGIES && slot < 0; j++) {
      if (used[j]) continue;
      const int ty = j < nl ? slot_list_type(list, j) : -2;
      if (hold) { if (ty == LETF_ST_BENCHMARK || ty == -2) slot = j; continue; }
      if (ty != st.type) continue;
      const int as = slot_list_asset(list, j);
      if (as >= 0 && st.asset != as) continue;
      if (st.type == LETF_ST_COMPOSITE && st.asset2 != 4) continue;
      slot = j;
    }
    if (slot < 0) return false;
    used[slot] = true; slot_of[q] = slot;
    if (!hold) *has_rules = true;
  }
  return true;
}
// the parameter block of a routed launch: the typed list's slots with the requested strategies in theirs (the others are
// placeholders of the slot's type, switched off by route_slots), the list's series as the asset mask
inline LetfParams routed_params(const LetfParams& s, int list, const int* slot_of, int n_slots, int mask) {
  LetfParams r = s;
  r.n_strategies = n_slots; r.asset_mask = mask;
  const int nl = slot_list_count(list);
  for (int j = 0; j < n_slots; j++) {
    LetfStrategyParams& st = r.strategies[j];
    memset(&st, 0, sizeof(st));
    const int ty = j < nl ? slot_list_type(list, j) : -2;
    st.type = ty >= 0 ? ty : LETF_ST_BENCHMARK;
    const int as = j < nl ? slot_list_asset(list, j) : -1;
    st.asset = as >= 0 ? as : 0; st.asset2 = 4;
  }
  for (int q = 0; q < s.n_strategies; q++) r.strategies[slot_of[q]] = s.strategies[q];
  return r;
}
template <typename P>
inline void route_slots(P& p, const LetfParams& s, const int* slot_of, int n_slots) {
  for (int j = 0; j < n_slots; j++) p.strat[j].route = -1;
  for (int q = 0; q < s.n_strategies; q++) p.strat[slot_of[q]].route = q + 1;
}
inline int routed_slot_count(const LetfParams& s, int list, const int* slot_of) {
  int n = slot_list_count(list);
  for (int q = 0; q < s.n_strategies; q++) if (slot_of[q] + 1 > n) n = slot_of[q] + 1;
  return n;
}

// v[a] for a run-time asset index without forcing the array into local memory: a select chain
// (in the unrolled builds `a` is loop-invariant and uniform, so this is a handful of predicated moves)
template <typename R> LETF_HD R pick6(const R (&v)[LETF_N_ASSETS], int a) {
  R r = v[0];
#pragma unroll
  for (int j = 1; j < LETF_N_ASSETS; j++) r = a == j ? v[j] : r;
  return r;
}

// ---- record staging of the rule kernels (Trial::run, MARKET == kMarketFSeries) -----------------------------------------------
constexpr int kStage = 4;                                     // days in flight
template <int NF> LETF_HD constexpr int kRecValues() { return NF + 9; }      // fields + the 9 retired-value lags
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float* letf_stage_smem() { extern __shared__ float letf_stage_buf[]; return letf_stage_buf; }
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#endif

#define TM_HAS(t) ((TM >> (t)) & 1u)
constexpr int kMarketSeries = 3;
constexpr int kMarketFSeries = 5;     // fp32 day-major records written by the pair producer (letf_jfast.cuh EMIT): split strategy pipeline
constexpr uint32_t kVolTypes = (1u << LETF_ST_VOL_TARGETING) | (1u << LETF_ST_ADAPTIVE_VOL) | (1u << LETF_ST_DOWNSIDE_VOL) |
                               (1u << LETF_ST_CONVEX_VOL) | (1u << LETF_ST_VOL_VELOCITY) | (1u << LETF_ST_VOL_MOM) |
                               (1u << LETF_ST_SKEW_CONVEX) | (1u << LETF_ST_META_ENSEMBLE) | (1u << LETF_ST_REGIME_ASYMMETRIC) |
                               (1u << LETF_ST_SKEW_KELLY) | (1u << LETF_ST_MOM_VOL_CONV) | (1u << LETF_ST_CONVICTION);

// ------------------------------------------------------------------------------------------
// max-drawdown tracker.  fp64: the reference's formula on the equity curve (engine.py:581-586).
// fp32: q = W / running_max evolves as q <- min(1, q * (1 + ret)), no division, no fp64.
template <typename R> struct Drawdown;
template <> struct Drawdown<double> {
  double rm, mn;
  LETF_HD void init(double w0) { rm = w0; mn = 0.0; }
  LETF_HD void step(double w, double) {
    if (w > rm) rm = w;
    double dd = (w - rm) / (rm > 0 ? rm : 1.0);
    if (dd < mn) mn = dd;
  }
  LETF_HD double value() const { return ::fabs(mn); }
  LETF_HD void init_open() { rm = 0; mn = 0; }
  LETF_HD void step_ret(double) {}
};
template <> struct Drawdown<float> {
  float q, qmin;
  LETF_HD void init(double) { q = 1.0f; qmin = 1.0f; }
  // branch-free start: the first step maps q = +inf to min(1, inf * g) = 1 for any g > 0
  LETF_HD void init_open() { q = INFINITY; qmin = 1.0f; }
  LETF_HD void step(double, float growth) {
    q = fminf(1.0f, q * growth);
    qmin = fminf(qmin, q);
  }
  // same with growth = 1 + ret folded into one FMA
  LETF_HD void step_ret(float ret) {
    q = fminf(1.0f, fmaf(q, ret, q));
    qmin = fminf(qmin, q);
  }
  LETF_HD double value() const { return 1.0 - (double)qmin; }   // in fp64: drawdowns within 1e-7 of 100 % (30-year 3x paths) keep their digits
};

// ------------------------------------------------------------------------------------------
template <typename R, int NS, int NA, int MARKET, bool TAPEIO, uint32_t TM, int LIST = 0, int AM = 0>
struct Trial {
  // LIST: typed slot list (0 = dynamic); AM: compile-time asset mask (0 = p.act_on at run time)
  typedef Mth<R> M;
  static constexpr bool EXACT = sizeof(R) == 8;
  static constexpr bool FSER = MARKET == kMarketFSeries;
  static constexpr bool ROUTED = slot_list_routed(LIST);     // slots can be switched off / write another output row (DevStrategy::route)
  // record mode on a frame with its own Cash_Ret column (KArgs::rec_cash): the routed and dynamic-slot rule kernels only - the
  // typed kernels of the benchmarked sweeps stay as they are (the test + select cost c4 1.4 %, measured)
  static constexpr bool CASHCOL = MARKET == kMarketFSeries && (LIST == 0 || slot_list_routed(LIST));
  static constexpr bool SERIES = MARKET == kMarketSeries || FSER;
  static constexpr int NF = FSER ? __builtin_popcount((unsigned)AM & 0x3Fu) + 2 : 1;      // record fields: active returns, VIX, SPY price
  static constexpr int kFSpy = __builtin_popcount((unsigned)AM & 0x0Fu);                 // record field of the SPY return (asset 4)
  static_assert(!FSER || (AM != 0 && ((AM >> 4) & 1) && (LIST == 0 || (AM & 1)) && !TAPEIO && !EXACT), "record mode: fp32 builds with a compile-time asset mask that holds SPY (typed lists: signal asset TQQQ = field 0)");
  static constexpr bool JOURNAL = TAPEIO;   // journal counters only in the parity/tape build
  static constexpr bool KEYREF = NA < kAssets; // calling convention of the non-inlined sampler retry loop (letf_math.cuh)
  static constexpr bool HAS_VOL = (TM & kVolTypes) != 0;
  static constexpr bool HAS_ACTIVE = (TM & ~((1u << LETF_ST_BENCHMARK) | (1u << LETF_ST_BENCHMARK_TLT))) != 0;
  static constexpr bool HAS_SMA200 = TM_HAS(LETF_ST_SMA) || TM_HAS(LETF_ST_SMA_BAND) || TM_HAS(LETF_ST_COMPOSITE) ||
                                     TM_HAS(LETF_ST_ADAPTIVE_VOL) || TM_HAS(LETF_ST_CONVEX_VOL) || TM_HAS(LETF_ST_META_ENSEMBLE);
  static constexpr bool HAS_SMA100 = TM_HAS(LETF_ST_CONVICTION);
  static constexpr bool HAS_RSI = TM_HAS(LETF_ST_COMPOSITE);
  static constexpr bool HAS_MOM = TM_HAS(LETF_ST_MOM_VOL_CONV) || TM_HAS(LETF_ST_CONVICTION);
  static constexpr bool HAS_V5 = TM_HAS(LETF_ST_VOL_VELOCITY) || TM_HAS(LETF_ST_META_ENSEMBLE) || TM_HAS(LETF_ST_REGIME_ASYMMETRIC) || TM_HAS(LETF_ST_CONVICTION);
  static constexpr bool HAS_V10 = TM_HAS(LETF_ST_MOM_VOL_CONV);
  static constexpr bool HAS_V20 = TM_HAS(LETF_ST_VOL_TARGETING) || TM_HAS(LETF_ST_ADAPTIVE_VOL) || TM_HAS(LETF_ST_CONVEX_VOL) ||
                                  TM_HAS(LETF_ST_VOL_VELOCITY) || TM_HAS(LETF_ST_VOL_MOM) || TM_HAS(LETF_ST_SKEW_CONVEX) ||
                                  TM_HAS(LETF_ST_META_ENSEMBLE) || TM_HAS(LETF_ST_SKEW_KELLY);
  static constexpr bool HAS_V60 = TM_HAS(LETF_ST_REGIME_ASYMMETRIC) || TM_HAS(LETF_ST_MOM_VOL_CONV) || TM_HAS(LETF_ST_CONVICTION);
  static constexpr bool HAS_DV20 = TM_HAS(LETF_ST_DOWNSIDE_VOL) || TM_HAS(LETF_ST_META_ENSEMBLE) || TM_HAS(LETF_ST_SKEW_KELLY) || TM_HAS(LETF_ST_CONVICTION);
  static constexpr bool HAS_SKEW = TM_HAS(LETF_ST_SKEW_CONVEX) || TM_HAS(LETF_ST_SKEW_KELLY);
  static constexpr bool HAS_MEAN60 = TM_HAS(LETF_ST_SKEW_KELLY);
  static constexpr bool HAS_VOLMOM = TM_HAS(LETF_ST_VOL_MOM);
  static constexpr bool NEED_RINGP = HAS_SMA200 || HAS_SMA100 || HAS_RSI;
  static constexpr bool NEED_RINGS = HAS_MOM;
  static constexpr bool NEED_RINGA = HAS_VOL;

  const KArgs<R, NS, NA>& ka;
  const DevParams<R, NS, NA>& p;
  int64_t k, col, lane, tcol, nt;
  bool use_tape, rec, anti;
  RngKey key;
  uint32_t flags;

  LETF_HD Trial(const KArgs<R, NS, NA>& a, int64_t trial, int64_t column, int64_t ln)
      : ka(a), p(a.p), k(trial), col(column), lane(ln) {
    int64_t sim_id = p.antithetic ? (k & ~(int64_t)1) : k;
    anti = p.antithetic && (k & 1);
    tcol = (p.antithetic ? (k >> 1) : k) - a.tape_first;
    nt = a.tape.n_tapes;
    use_tape = TAPEIO && a.rng_mode == LETF_RNG_TAPE;
    rec = TAPEIO && a.rng_mode == LETF_RNG_PHILOX_RECORD;
    key.k0 = (uint32_t)a.seed; key.k1 = (uint32_t)(a.seed >> 32);
    key.t0 = (uint32_t)sim_id; key.t1 = (uint32_t)((uint64_t)sim_id >> 32);
    flags = 0;
  }

  // ---- tape helpers -------------------------------------------------------------------
  LETF_HD R tp_get(const double* a, int64_t row) const { return (R)a[row * nt + tcol]; }
  LETF_HD void tp_put(double* a, int64_t row, R v) const { a[row * nt + tcol] = (double)v; }

  // ---- regime path ---------------------------------------------------------------------
  struct RegimeGen { int cur, left, ks; };
  int start_reg = 0;      // random_start.py:31 / :106 ('regime_only', 'historical_anchor'); 0 otherwise
  LETF_HD void regime_init(RegimeGen& g) const { g.cur = start_reg; g.left = 0; g.ks = 0; }
  LETF_HD int draw_spell(int ks, int cur) {
    const int G = p.T + p.buffer;
    int spell;
    if (use_tape) {
      if (ks >= ka.tape.max_spells) { flags |= LETF_FLAG_SPELL_OVERFLOW; spell = G; }
      else spell = ka.tape.spells[(int64_t)ks * nt + tcol];
    } else {
      U4 b = rng_block(key, (uint32_t)ks, STREAM_SPELL);
      if (p.n_dur[cur] > 0) spell = p.dur[cur][mulhi32(b.y, (uint32_t)p.n_dur[cur])];
      else {
        R q = M::div_(M::log_(M::u01(b.x)), p.log_pstay[cur]);   // inverse CDF of the geometric dwell time
        spell = q < (R)G ? (int)q + 1 : G;
      }
      if (rec && ks < ka.tape.max_spells) ka.tape.spells[(int64_t)ks * nt + tcol] = spell;
    }
    return spell < 1 ? 1 : spell;
  }
  LETF_HD int regime_next(RegimeGen& g) {
    if (MARKET == LETF_MARKET_NORMAL) return 0;
    if (g.left == 0) {
      if (g.ks > 0 && p.flip[g.cur]) g.cur ^= 1;
      g.left = draw_spell(g.ks, g.cur);
      g.ks++;
    }
    g.left--;
    return g.cur;
  }

  // ---- market generators -----------------------------------------------------------------
  struct JointState { R h[3], q[6], pz[3]; int prev_reg; };

  // one generated day of the raw (pre moment-matching) underlying returns x[3] = SPY, QQQ, TLT.
  // Also returns the two "phase-2" normals of the day (VIX noise, IRX noise) when WANT_AUX.
  template <bool WANT_X, bool WANT_AUX>
  LETF_HD void market_day(int g, int reg, JointState& js, R x[3], R& vix_n, R& irx_n) {
    market_day<WANT_X, WANT_AUX>(g, reg, js, x, vix_n, irx_n, p.joint[reg]);
  }
  // jp: the joint-t constants of `reg` (pass 1 of the two-pass builds keeps them in registers between regime switches)
  template <bool WANT_X, bool WANT_AUX>
  LETF_HD void market_day(int g, int reg, JointState& js, R x[3], R& vix_n, R& irx_n, const DevJoint<R>& jp) {
    R z0 = 0, z1 = 0, z2 = 0, nchi = 0;
    uint32_t cw0 = 0, cw1 = 0;
    if (use_tape) {
      if (WANT_X) { z0 = tp_get(ka.tape.x, (int64_t)g * 3 + 0); z1 = tp_get(ka.tape.x, (int64_t)g * 3 + 1); z2 = tp_get(ka.tape.x, (int64_t)g * 3 + 2); }
      if (WANT_AUX && MARKET != LETF_MARKET_NORMAL) { vix_n = tp_get(ka.tape.vix_n, g); irx_n = tp_get(ka.tape.irx_n, g); }
    } else {
      if (WANT_X && MARKET == LETF_MARKET_NORMAL) {
        // iid-normal shapes: the SPY / QQQ shocks of days 2j and 2j+1 share one block (words x,y / z,w of
        // (j, STREAM_MKT_PAIR)); the third shock has its own block and is drawn only when a TLT-based
        // series is computed (letf_fast.cuh consumes the same layout two days per iteration)
        U4 bp = rng_block(key, (uint32_t)g >> 1, STREAM_MKT_PAIR);
        if (g & 1) M::box_muller(bp.z, bp.w, z0, z1); else M::box_muller(bp.x, bp.y, z0, z1);
        if (TAPEIO || p.need_tlt) {
          U4 b0 = rng_block(key, (uint32_t)g, STREAM_MKT0);
          M::box_muller(b0.z, b0.w, z2, nchi);
        }
      }
      if (WANT_X && MARKET == LETF_MARKET_JOINT_T) {
        U4 b0 = rng_block(key, (uint32_t)g, STREAM_MKT0);
        M::box_muller(b0.x, b0.y, z0, z1);
        M::box_muller(b0.z, b0.w, z2, nchi);
      }
      if (WANT_X && MARKET == LETF_MARKET_PARAM_T) {
        for (int c = 0; c < 3; c++) {
          U4 b = rng_block(key, (uint32_t)g, STREAM_PT0 + c);
          R t = student_t_block<R, KEYREF>(b, p.pt_df, p.pt_fast_k, key, (uint32_t)g, STREAM_PT0 + c);
          if (c == 0) z0 = t; else if (c == 1) z1 = t; else z2 = t;
        }
      }
      if (MARKET != LETF_MARKET_NORMAL && (WANT_AUX || (WANT_X && MARKET == LETF_MARKET_JOINT_T))) {
        U4 b1 = rng_block(key, (uint32_t)g, STREAM_MKT1);
        if (WANT_AUX) M::box_muller(b1.x, b1.y, vix_n, irx_n);
        cw0 = b1.z; cw1 = b1.w;
      }
      if (rec) {
        if (WANT_X) { tp_put(ka.tape.x, (int64_t)g * 3 + 0, z0); tp_put(ka.tape.x, (int64_t)g * 3 + 1, z1); tp_put(ka.tape.x, (int64_t)g * 3 + 2, z2); }
        if (WANT_AUX && MARKET != LETF_MARKET_NORMAL) { tp_put(ka.tape.vix_n, g, vix_n); tp_put(ka.tape.irx_n, g, irx_n); }
      }
    }
    if (!WANT_X) return;
    if (MARKET == LETF_MARKET_NORMAL) {
      x[0] = p.n_mu[0] + p.n_ch[0] * z0;
      x[1] = p.n_mu[1] + (p.n_ch[1] * z0 + p.n_ch[2] * z1);
      x[2] = p.n_mu[2] + (p.n_ch[3] * z0 + p.n_ch[4] * z1 + p.n_ch[5] * z2);
    } else if (MARKET == LETF_MARKET_PARAM_T) {
      const R* C = p.pt_chol[reg];
      R xa = C[0] * z0 + C[1] * z1 + C[2] * z2;
      R xb = C[3] * z0 + C[4] * z1 + C[5] * z2;
      R xc = C[6] * z0 + C[7] * z1 + C[8] * z2;
      x[0] = p.reg_mean[reg] + p.reg_std[reg] * xa;
      x[1] = (R)1.15 * x[0] + (R)0.006 * xb;
      x[2] = (R)-0.12 * x[0] + (R)0.004 * xc;
    } else {
      // chi-square mixing variable (calibration.py:565)
      R u;
      if (use_tape) u = tp_get(ka.tape.chi2, g);
      else {
        u = jp.nu_fast_k ? chi2_int_fast<R>(jp.nu_fast_k, cw0, cw1, nchi)
                         : chi2_marsaglia_tsang<R, KEYREF>(jp.nu, nchi, cw0, key, (uint32_t)g, STREAM_MKT1);
        if (rec) tp_put(ka.tape.chi2, g, u);
      }
      // GARCH / DCC recursions (calibration.py:532-548)
      if (g == 0) {
        for (int a = 0; a < 3; a++) js.h[a] = jp.lv[a];
        for (int a = 0; a < 6; a++) js.q[a] = jp.rb[a];
      } else if (reg != js.prev_reg) {
        const R bl = (R)0.10;
        for (int a = 0; a < 3; a++) js.h[a] = bl * jp.lv[a] + ((R)1 - bl) * js.h[a];
        for (int a = 0; a < 6; a++) js.q[a] = bl * jp.rb[a] + ((R)1 - bl) * js.q[a];
      } else if (!EXACT) {
        // same recursions with the per-regime constants folded on the host: 3 operations per element
        for (int a = 0; a < 3; a++) js.h[a] = rmax(M::fma_(js.h[a], M::fma_(jp.alpha, js.pz[a] * js.pz[a], jp.beta), jp.c0[a]), (R)1e-10);
        js.q[0] = M::fma_(jp.dcc_b, js.q[0], M::fma_(jp.dcc_a, js.pz[0] * js.pz[0], jp.c1[0]));
        js.q[1] = M::fma_(jp.dcc_b, js.q[1], M::fma_(jp.dcc_a, js.pz[1] * js.pz[0], jp.c1[1]));
        js.q[2] = M::fma_(jp.dcc_b, js.q[2], M::fma_(jp.dcc_a, js.pz[1] * js.pz[1], jp.c1[2]));
        js.q[3] = M::fma_(jp.dcc_b, js.q[3], M::fma_(jp.dcc_a, js.pz[2] * js.pz[0], jp.c1[3]));
        js.q[4] = M::fma_(jp.dcc_b, js.q[4], M::fma_(jp.dcc_a, js.pz[2] * js.pz[1], jp.c1[4]));
        js.q[5] = M::fma_(jp.dcc_b, js.q[5], M::fma_(jp.dcc_a, js.pz[2] * js.pz[2], jp.c1[5]));
      } else {
        for (int a = 0; a < 3; a++) {
          R hn = jp.one_m_ab * jp.lv[a] + jp.alpha * (js.pz[a] * js.pz[a]) * js.h[a] + jp.beta * js.h[a];
          js.h[a] = hn < (R)1e-10 ? (R)1e-10 : hn;
        }
        js.q[0] = jp.one_m_dcc * jp.rb[0] + jp.dcc_a * (js.pz[0] * js.pz[0]) + jp.dcc_b * js.q[0];
        js.q[1] = jp.one_m_dcc * jp.rb[1] + jp.dcc_a * (js.pz[1] * js.pz[0]) + jp.dcc_b * js.q[1];
        js.q[2] = jp.one_m_dcc * jp.rb[2] + jp.dcc_a * (js.pz[1] * js.pz[1]) + jp.dcc_b * js.q[2];
        js.q[3] = jp.one_m_dcc * jp.rb[3] + jp.dcc_a * (js.pz[2] * js.pz[0]) + jp.dcc_b * js.q[3];
        js.q[4] = jp.one_m_dcc * jp.rb[4] + jp.dcc_a * (js.pz[2] * js.pz[1]) + jp.dcc_b * js.q[4];
        js.q[5] = jp.one_m_dcc * jp.rb[5] + jp.dcc_a * (js.pz[2] * js.pz[2]) + jp.dcc_b * js.q[5];
      }
      js.prev_reg = reg;
      R zz[3];
      R ih[3] = {(R)0, (R)0, (R)0};      // fp32: 1 / sqrt(max(h, 1e-10)), shared by the factor and by prev_z
      R sh0, sh1, sh2;
      if (EXACT) { sh0 = M::sqrt_(js.h[0]); sh1 = M::sqrt_(js.h[1]); sh2 = M::sqrt_(js.h[2]); }
      else {
        for (int a = 0; a < 3; a++) ih[a] = M::rsqrt_(rmax(js.h[a], (R)1e-10));
        sh0 = js.h[0] * ih[0]; sh1 = js.h[1] * ih[1]; sh2 = js.h[2] * ih[2];
      }
      if (EXACT) {
        // cov_t = outer(sqrt h) * Q / outer(d, d), then its Cholesky factor (calibration.py:556-561)
        R d0 = M::sqrt_(rmax(js.q[0], (R)1e-12)), d1 = M::sqrt_(rmax(js.q[2], (R)1e-12)), d2 = M::sqrt_(rmax(js.q[5], (R)1e-12));
        R c00 = (sh0 * sh0) * (js.q[0] / (d0 * d0)), c10 = (sh1 * sh0) * (js.q[1] / (d1 * d0));
        R c11 = (sh1 * sh1) * (js.q[2] / (d1 * d1)), c20 = (sh2 * sh0) * (js.q[3] / (d2 * d0));
        R c21 = (sh2 * sh1) * (js.q[4] / (d2 * d1)), c22 = (sh2 * sh2) * (js.q[5] / (d2 * d2));
        R l00 = M::sqrt_(c00), l10 = c10 / l00, l20 = c20 / l00;
        R l11 = M::sqrt_(c11 - l10 * l10), l21 = (c21 - l20 * l10) / l11;
        R l22 = M::sqrt_(c22 - l20 * l20 - l21 * l21);
        zz[0] = l00 * z0;
        zz[1] = l10 * z0 + l11 * z1;
        zz[2] = l20 * z0 + l21 * z1 + l22 * z2;
      } else {
        // same factor written as diag(sqrt h) * chol(R_t) with R_t unit-diagonal: fewer operations
        R i0 = M::rsqrt_(rmax(js.q[0], (R)1e-12)), i1 = M::rsqrt_(rmax(js.q[2], (R)1e-12)), i2 = M::rsqrt_(rmax(js.q[5], (R)1e-12));
        R r10 = js.q[1] * i1 * i0, r20 = js.q[3] * i2 * i0, r21 = js.q[4] * i2 * i1;
        const R s11 = rmax((R)1 - r10 * r10, (R)0);
        const R il11 = M::rsqrt_(rmax(s11, (R)1e-30));
        R l11 = s11 * il11;                                        // sqrt(s11)
        R l21 = (r21 - r20 * r10) * il11;
        R l22 = M::sqrt_(rmax((R)1 - r20 * r20 - l21 * l21, (R)0));
        zz[0] = sh0 * z0;
        zz[1] = sh1 * (r10 * z0 + l11 * z1);
        zz[2] = sh2 * (r20 * z0 + l21 * z1 + l22 * z2);
      }
      if (anti) { zz[0] = -zz[0]; zz[1] = -zz[1]; zz[2] = -zz[2]; }
      R scale = EXACT ? M::sqrt_(M::div_(jp.nu, rmax(u, (R)1e-12))) : jp.sqrt_nu * M::rsqrt_(rmax(u, (R)1e-12));
      for (int a = 0; a < 3; a++) {
        R xx = jp.mu[a] + zz[a] * scale;
        R xc = clipr(xx, (R)-0.95, (R)4.0);
        js.pz[a] = EXACT ? (xc - jp.mu[a]) / M::sqrt_(rmax(js.h[a], (R)1e-10)) : (xc - jp.mu[a]) * ih[a];
        x[a] = xc;
      }
    }
  }

  // ---- legacy_hybrid block bootstrap (bootstrap.py:310-458 + engine.py:50-80) --------------------
  // one block draw from a (pool, regime): momentum continuation picks the same-sign candidates with probability
  // `bias` (bootstrap.py:224-229); pool B is then weighted by a Gaussian kernel around the economy block's SPY
  // return (:232-259, numpy's choice = cdf.searchsorted(u, 'right')), pool A is uniform (:262).  Returns the block's
  // first history row.
  LETF_HD int boot_pick(int pool, int regime, bool have_last, R last_ret, R u_mom, R u_idx, bool use_target, R target) const {
    const DevBoot<R>& b = p.boot;
    int kind = 0;
    if (have_last) {
      const int want = last_ret >= (R)0 ? 1 : 2;
      if (b.cnt[pool][regime][want] > 0 && b.bias[regime] > (R)0 && u_mom < b.bias[regime]) kind = want;
    }
    const int nc = b.cnt[pool][regime][kind];
    const int* cs = b.cstart + b.off[pool][regime][kind];
    const R* cr = b.cret + b.off[pool][regime][kind];
    int idx = 0;
    if (use_target && nc > 5) {
      const R sigma = b.sigma[regime][kind];
      R W = 0;
      for (int i = 0; i < nc; i++) { const R d = M::div_(cr[i] - target, sigma); W += M::exp_((R)-0.5 * (d * d)); }
      R cum = 0;
      for (int i = 0; i < nc; i++) {
        const R d = M::div_(cr[i] - target, sigma);
        cum += M::div_(M::exp_((R)-0.5 * (d * d)), W);
        if (cum <= u_idx) idx = i + 1;
      }
    } else {
      idx = (int)M::floor_(u_idx * (R)nc);
    }
    if (idx > nc - 1) idx = nc - 1;
    return cs[idx];
  }

  // pass 1: stitch the blocks into the per-thread scratch column and return the three column means
  template <typename X4>
  LETF_HD void boot_stitch(X4* xs, R (&mean)[3]) {
    const DevBoot<R>& b = p.boot;
    const int G = p.T + p.buffer;
    RegimeGen rgl; regime_init(rgl);
    R sum[3] = {0, 0, 0}, cmp[3] = {0, 0, 0};
    int cur = 0, kb = 0;
    bool have_last = false;
    R last_a = 0, last_b = 0;
    while (cur < G) {
      int len;
      R u[6];
      if (use_tape) {
        if (kb >= ka.tape.max_blocks) { flags |= LETF_FLAG_SPELL_OVERFLOW; len = G; for (int s = 0; s < 6; s++) u[s] = 0; }
        else {
          len = ka.tape.boot_len[(int64_t)kb * nt + tcol];
          for (int s = 0; s < 6; s++) u[s] = tp_get(ka.tape.boot_u, (int64_t)kb * 6 + s);
        }
      } else {
        const U4 wa = rng_block(key, (uint32_t)kb, STREAM_BOOT_A), wb = rng_block(key, (uint32_t)kb, STREAM_BOOT_B);
        const R q = M::div_(M::log_(M::u01(wa.x)), b.log1mp);         // geometric(1 / block_mean), bootstrap.py:304
        len = q < (R)1e6 ? (int)q + 1 : 1000000;
        u[0] = M::u01(wa.y); u[1] = M::u01(wa.z); u[2] = M::u01(wa.w);
        u[3] = M::u01(wb.x); u[4] = M::u01(wb.y); u[5] = M::u01(wb.z);
        if (rec && kb < ka.tape.max_blocks) {
          ka.tape.boot_len[(int64_t)kb * nt + tcol] = len;
          for (int s = 0; s < 6; s++) tp_put(ka.tape.boot_u, (int64_t)kb * 6 + s, u[s]);
        }
      }
      len = len < b.block_min ? b.block_min : (len > b.block_max ? b.block_max : len);
      if (len > G - cur) len = G - cur;
      int ones = 0;
      for (int j = 0; j < len; j++) ones += regime_next(rgl);
      const int regime = 2 * ones > len ? 1 : 0;                      // int(np.median(block_regime_path)), bootstrap.py:340
      const int span = b.block_size - len + 1;                        // sub-section starts (bootstrap.py:350-352)
      // an empty pool is served by a synthetic block (bootstrap.py:214-222, 273-300): Student-t draws times a regime
      // volatility for SPY, QQQ = 1.25 SPY, TLT = -0.25 SPY; no momentum coin, no block index, same sub-section rule
      const bool syn_a = b.cnt[0][regime][0] == 0, syn_b = b.cnt[1][regime][0] == 0;
      const R syn_std = regime == 0 ? (R)0.011 : (R)0.035;
      auto syn = [&](int q, int row) -> R {                           // row of block kb's synthetic block for pool q
        R t;
        const int64_t slot = ((int64_t)kb * 2 + q) * ka.tape.boot_block_size + row;
        if (use_tape) t = ka.tape.boot_syn ? tp_get(ka.tape.boot_syn, slot) : (R)0;
        else {
          const uint32_t ctr = (uint32_t)kb * (uint32_t)b.block_size + (uint32_t)row;
          const U4 w = rng_block(key, ctr, STREAM_BOOT_SYN + q);
          t = student_t_block<R, KEYREF>(w, p.pt_df, p.pt_fast_k, key, ctr, STREAM_BOOT_SYN + q);
          if (rec && ka.tape.boot_syn && kb < ka.tape.max_blocks) tp_put(ka.tape.boot_syn, slot, t);
        }
        return t * syn_std;
      };
      int row_a = syn_a ? 0 : boot_pick(0, regime, have_last, last_a, u[0], u[1], false, (R)0);
      if (len < b.block_size) { int o = (int)M::floor_(u[2] * (R)span); row_a += o > span - 1 ? span - 1 : o; }
      const R* hs = b.hist[0] + (syn_a ? 0 : row_a);
      R prod_a = (R)1;
      if (!syn_a) { for (int j = 0; j < len; j++) prod_a *= (R)1 + hs[j]; }          // (the common case stays a plain gather loop)
      else { for (int j = 0; j < len; j++) prod_a *= (R)1 + syn(0, row_a + j); }
      const R spy_ret = prod_a - (R)1;
      int row_b = syn_b ? 0 : boot_pick(1, regime, have_last, last_b, u[3], u[4], true, spy_ret);
      if (len < b.block_size) { int o = (int)M::floor_(u[5] * (R)span); row_b += o > span - 1 ? span - 1 : o; }
      const R* hq = b.hist[1] + (syn_b ? 0 : row_b);
      const R* ht = b.hist[2] + (syn_b ? 0 : row_b);
      R prod_b = (R)1;
      auto put = [&](int j, R vx, R vy, R vz) {
        X4 v; v.x = vx; v.y = vy; v.z = vz; v.w = (R)0;
        xs[cur + j] = v;
        prod_b *= (R)1 + vy;
        const R xv[3] = {vx, vy, vz};
        for (int a = 0; a < 3; a++) { R y = xv[a] - cmp[a], t = sum[a] + y; cmp[a] = (t - sum[a]) - y; sum[a] = t; }
      };
      if (!syn_a && !syn_b) { for (int j = 0; j < len; j++) put(j, hs[j], hq[j], ht[j]); }
      else {
        for (int j = 0; j < len; j++) {
          const R vx = syn_a ? syn(0, row_a + j) : hs[j];
          if (syn_b) { const R sb = syn(1, row_b + j); put(j, vx, sb * (R)1.25, -sb * (R)0.25); }
          else put(j, vx, hq[j], ht[j]);
        }
      }
      last_a = spy_ret; last_b = prod_b - (R)1; have_last = true;
      cur += len; kb++;
    }
    for (int a = 0; a < 3; a++) mean[a] = G > 0 ? sum[a] / (R)G : (R)0;
  }

  // pass 2, one generated day: both correlated Student-t blends on the stitched returns, and the day's VIX / IRX normals
  LETF_HD void boot_day(int g, int reg, R b0, R b1, R b2, const R (&mean)[3], R x[3], R& vix_n, R& irx_n) {
    const auto& b = p.boot;
    R z1[3], z2[3];
    if (use_tape) {
      for (int c = 0; c < 3; c++) { z1[c] = tp_get(ka.tape.boot_t, (int64_t)g * 3 + c); z2[c] = tp_get(ka.tape.x, (int64_t)g * 3 + c); }
      vix_n = tp_get(ka.tape.vix_n, g); irx_n = tp_get(ka.tape.irx_n, g);
    } else {
      if (!EXACT && p.pt_fast_k == 5) {
        // cfg.STUDENT_T_DF's default: the six variates straight-line, their Philox blocks three at a time with
        // interleaved rounds (same counters and arithmetic as the general branch below)
        const uint32_t s1[3] = {STREAM_BN0, STREAM_BN0 + 1, STREAM_BN0 + 2}, s2[3] = {STREAM_PT0, STREAM_PT0 + 1, STREAM_PT0 + 2};
        U4 w1[3], w2[3];
        rng_blocks<3>(key, (uint32_t)g, s1, w1);
        rng_blocks<3>(key, (uint32_t)g, s2, w2);
#pragma unroll
        for (int c = 0; c < 3; c++) { z1[c] = student_t5<R>(w1[c]); z2[c] = student_t5<R>(w2[c]); }
      } else
      for (int c = 0; c < 3; c++) {
        const U4 w1 = rng_block(key, (uint32_t)g, STREAM_BN0 + c);
        z1[c] = student_t_block<R, KEYREF>(w1, p.pt_df, p.pt_fast_k, key, (uint32_t)g, STREAM_BN0 + c);
        const U4 w2 = rng_block(key, (uint32_t)g, STREAM_PT0 + c);
        z2[c] = student_t_block<R, KEYREF>(w2, p.pt_df, p.pt_fast_k, key, (uint32_t)g, STREAM_PT0 + c);
      }
      const U4 w = rng_block(key, (uint32_t)g, STREAM_MKT1);
      M::box_muller(w.x, w.y, vix_n, irx_n);
      if (rec) {
        for (int c = 0; c < 3; c++) { tp_put(ka.tape.boot_t, (int64_t)g * 3 + c, z1[c]); tp_put(ka.tape.x, (int64_t)g * 3 + c, z2[c]); }
        tp_put(ka.tape.vix_n, g, vix_n); tp_put(ka.tape.irx_n, g, irx_n);
      }
    }
    if (b.nw1 > (R)0) {                                               // bootstrap.py:386-449
      // (indexed through the parameter block, not through a pointer into it: constant-bank loads, no generic LD)
#define C(i) b.ch1[reg][i]
      const R c0 = C(0) * z1[0] + C(1) * z1[1] + C(2) * z1[2];
      const R c1 = C(3) * z1[0] + C(4) * z1[1] + C(5) * z1[2];
      const R c2 = C(6) * z1[0] + C(7) * z1[1] + C(8) * z1[2];
#undef C
      const R s0 = reg == 0 ? (R)0.007 : (R)0.022, s1 = s0 * (R)1.35, s2 = s0 * (R)0.5;
      b0 = b.w1 * b0 + b.nw1 * (c0 * s0 + mean[0]);
      b1 = b.w1 * b1 + b.nw1 * (c1 * s1 + mean[1]);
      b2 = b.w1 * b2 + b.nw1 * (c2 * s2 + mean[2]);
    }
    {                                                                 // engine.py:57-79
#define C(i) b.ch2[reg][i]
      const R x0 = C(0) * z2[0] + C(1) * z2[1] + C(2) * z2[2];
      const R x1 = C(3) * z2[0] + C(4) * z2[1] + C(5) * z2[2];
      const R x2 = C(6) * z2[0] + C(7) * z2[1] + C(8) * z2[2];
#undef C
      const R sd = p.reg_std[reg], qsd = (R)1.35 * sd, tsd = (R)0.55 * sd;
      const R n0 = b.noise_w2 * sd * (R)0.55 * x0, n1 = b.noise_w2 * qsd * (R)0.45 * x1, n2 = b.noise_w2 * tsd * (R)0.35 * x2;
      x[0] = b.w2 * b0 + b.omw2 * (b0 + n0);
      x[1] = b.w2 * b1 + b.omw2 * (b1 + n1);
      x[2] = b.w2 * b2 + b.omw2 * (b2 + n2);
    }
  }

  // ---- strategy bank ------------------------------------------------------------------------
  struct Signals {          // state "as of yesterday" (pandas .shift(1)) unless noted
    RollSum<R, EXACT> sma200, sma100, gain14, loss14, mom126, mean60;
    RollVar<R, EXACT> v5, v10, v20, v60, dv20;
    RollSkew<R, EXACT> sk60;
    R volhist[6];           // shifted 20-day vol series, newest first (pct_change(5))
    R P_prev, vix_prev;
    int cP, cS, cA;         // ring cursors = day index modulo ring depth
  };

  // fp32 throughput mode: plain sliding sums instead of pandas' compensated online kernels.  Returns have |mean| << std, so
  // the variance / skew from raw power sums is well conditioned; the sums are refreshed from the window's history every
  // kRefresh days (uniform: every thread is on the same day), which bounds the add/subtract drift.
  struct FastSig {
    R a1[4], a2[4];         // sum r, sum r^2 over the last 5 / 10 / 20 / 60 returns of the signal asset
    R a3;                   // sum r^3 over the last 60
    R d1, d2;               // sum min(r, 0) and its square over the last 20 (downside vol)
    R p200, p100;           // SPY price sums
    R g14, l14;             // RSI: gains / losses over the last 14 price differences
    R m126;                 // SPY return sum (momentum)
  };
  static constexpr int kRefresh = 128;

  struct Slot {
    double W;               // equity (always fp64: 7 560 multiplicative steps)
    Drawdown<R> dd;
    R aux;                  // prev_alloc | position | composite state
    int trades;
    R pv[kAssets];          // portfolio rule: dollar position per asset (strategy.py:155-241); unused elsewhere
  };
  struct SlotJ { double pos[3]; double vol; int cnt; };

  LETF_HD void jlog(SlotJ& j, int which, int s, int day, int asset, R new_alloc, double pv, R price) {
    // trade.py:20-73
    if (price <= (R)0) return;
    double target = (double)new_alloc * pv / (double)price;
    double ch = target - j.pos[which];
    if (::fabs(ch) < 0.001) return;
    double sh = ch > 0 ? ch : ::fabs(ch);
    if (TAPEIO && ka.log_count && (ka.trial_ids || ka.log_trial == k) &&
        ((ka.trial_ids && ka.log_strategies) ? ka.log_strategies[col] : ka.log_strategy) == s) {
      const int64_t row = ka.trial_ids ? col : 0;
      int n = ka.log_count[row];
      if (n < ka.log_cap) {
        const int64_t o = row * ka.log_cap + n;
        ka.log_day[o] = day; ka.log_asset[o] = asset; ka.log_action[o] = ch > 0 ? 1 : -1;
        ka.log_shares[o] = sh; ka.log_price[o] = (double)price;
      }
      ka.log_count[row] = n + 1;
    }
    j.cnt++;
    j.vol += sh * (double)price;
    j.pos[which] = target;
  }

  // pandas recomputes a rolling-variance window from scratch when a removal collapses it
  // (letf_rolling.cuh): re-add the w-1 surviving observations, oldest first, from the return ring
  // (cursor = slot today's return will be written to).  Rare: the window just became constant.
  // (static: a non-inlined MEMBER would take `this`, and once `this` escapes every parameter read of the day loop
  //  turns from a constant-bank load into a generic load through the saved reference)
  static LETF_HDN void revar_(RollVar<R, EXACT>& rv, const R* ring, int64_t nthreads, int cursor, int w, bool downside) {
    rv.init();
    for (int j = w - 1; j >= 1; j--) {
      int c = cursor - j; if (c < 0) c += kRingA;
      R v = ring[(int64_t)c * nthreads];
      rv.add(downside ? (v < (R)0 ? v : (R)0) : v);
    }
  }

  // run the whole trial ---------------------------------------------------------------------
  // (the loop-unswitched straight-line form of the lean buy&hold shapes lives in letf_fast.cuh)
  LETF_HD void run() {
    const int G = p.T + p.buffer;
    const int64_t nth = ka.nthreads;
    const uint32_t nth32 = (uint32_t)ka.nthreads;             // ring rows: depth * nthreads < 2^32 elements (checked by the host)
    auto roff = [nth32](int c) { return (uint32_t)c * nth32; };
    const R V = (R)15.874507866387544;   // sqrt(252)

    // ---- start offset (random_start.py:33-57)
    int so = 0;
    if (!SERIES && p.random_start) {
      if (use_tape) so = ka.tape.start_offset[tcol];
      else {
        U4 b = rng_block(key, 0, STREAM_START);
        so = p.min_offset + (int)mulhi32(b.x, (uint32_t)p.offset_range);
        if (rec) ka.tape.start_offset[tcol] = so;
      }
    }

    if (!SERIES && MARKET != LETF_MARKET_NORMAL && p.start_p1 > (R)0) {
      if (use_tape) start_reg = ka.tape.start_regime ? (ka.tape.start_regime[tcol] != 0) : 0;
      else {
        U4 b = rng_block(key, 0, STREAM_START);
        start_reg = M::u01(b.y) < p.start_p1 ? 1 : 0;
        if (rec && ka.tape.start_regime) ka.tape.start_regime[tcol] = start_reg;
      }
    }

    // ---- pass 1 (moment matching only): raw joint returns -> scratch, column means (calibration.py:579-588)
    struct alignas(4 * sizeof(R) > 16 ? 16 : 4 * sizeof(R)) X4 { R x, y, z, w; };
    X4* const xs = (X4*)ka.xscr + (int64_t)lane * (p.T + p.buffer);
    // Philox-only fp32 builds: pass 1 already holds the day's second market block (its z,w words feed the chi-square),
    // so it also turns the x,y words into the VIX / IRX normals and keeps them (w word of the day's vector + a second
    // per-thread column behind the vectors); pass 2 and the burn-in then draw no market block at all
    constexpr bool AUXSCR = !EXACT && !TAPEIO && MARKET == LETF_MARKET_JOINT_T && LIST != 2;   // (the 19-rule build has no registers to spare)
    // (the column is written four days at a time from registers: 4-byte stores would each dirty a sector of their own
    //  and come back from L2 as read-modify-write traffic; its per-thread stride is G rounded up to a multiple of 4)
    const int Gp = (p.T + p.buffer + 3) & ~3;
    R* const is = (R*)ka.xscr + (int64_t)4 * (p.T + p.buffer) * ka.nthreads + (int64_t)lane * Gp;
    R shift[3] = {(R)0, (R)0, (R)0};
    const bool two_pass = (MARKET == LETF_MARKET_JOINT_T && p.mm) || MARKET == kMarketBoot;
    const int T = SERIES ? ka.series.sim_days : p.T;
    JointState js;
    js.prev_reg = 0;
    for (int a = 0; a < 3; a++) { js.h[a] = 0; js.pz[a] = 0; }
    for (int a = 0; a < 6; a++) js.q[a] = 0;
    if (MARKET == kMarketBoot) {
      if (G > 0) boot_stitch(xs, shift);            // shift[] holds the column MEANS of the stitched returns here
    } else if (two_pass && G > 0) {
      RegimeGen rg; regime_init(rg);
      R sum[3] = {0, 0, 0}, cmp[3] = {0, 0, 0};
      X4 ib; ib.x = ib.y = ib.z = ib.w = (R)0;
      X4 vprev; vprev.x = vprev.y = vprev.z = vprev.w = (R)0;
      DevJoint<R> jpc = p.joint[0];            // regime-indexed constants cached in registers, reloaded on a switch
      int jreg = 0;
      for (int g = 0; g < G; g++) {
        int reg = regime_next(rg);
        R x[3], d0 = 0, d1 = 0;
        if (!EXACT && reg != jreg) { jpc = p.joint[reg]; jreg = reg; }
        if (EXACT) market_day<true, false>(g, reg, js, x, d0, d1);
        else if (AUXSCR) market_day<true, true>(g, reg, js, x, d0, d1, jpc);
        else market_day<true, false>(g, reg, js, x, d0, d1, jpc);
        X4 v; v.x = x[0]; v.y = x[1]; v.z = x[2]; v.w = d0;
        if (EXACT) xs[g] = v;
        else if (g & 1) { xs[g - 1] = vprev; xs[g] = v; }      // fp32: two days = one full 32-byte sector per store pair
        else vprev = v;
        if (AUXSCR) {
          ib.x = ib.y; ib.y = ib.z; ib.z = ib.w; ib.w = d1;
          if ((g & 3) == 3) *(X4*)(is + (g - 3)) = ib;
        }
        for (int a = 0; a < 3; a++) {
          R y = x[a] - cmp[a], t = sum[a] + y;      // Kahan: stays within an ulp of numpy's pairwise mean
          cmp[a] = (t - sum[a]) - y; sum[a] = t;
        }
      }
      if (!EXACT && (G & 1)) xs[G - 1] = vprev;
      if (AUXSCR && (G & 3)) {          // the last G % 4 days: shift them to the front of a final, padded vector
        for (int r = G & 3; r < 4; r++) { ib.x = ib.y; ib.y = ib.z; ib.z = ib.w; ib.w = (R)0; }
        *(X4*)(is + (G & ~3)) = ib;
      }
      for (int a = 0; a < 3; a++) shift[a] = (p.mm_theo[a] - sum[a] / (R)G) * (R)0.1;
    }

    // ---- pass 2 state
    RegimeGen rg; regime_init(rg);
    R vix = p.const_vix;
    // stress / tracking-error / price state
    R liq = 0, cred = 0;
    static_assert(NA == kAssets || (!HAS_ACTIVE && !SERIES), "compacted asset slots are for buy&hold-only builds");
    R te_prev[NA], bias[NA];
    double cp[NA];                   // cumulative growth of each ETF series (price / 100), always fp64
    Drawdown<R> ddA[NA], ddT;        // buy&hold statistics live on the asset, not on a strategy slot
    for (int a = 0; a < NA; a++) { te_prev[a] = 0; bias[a] = 0; cp[a] = 1.0; ddA[a].init_open(); }
    ddT.init_open();
    double cp_tlt = 1.0;
    R vmin = (R)INFINITY, vmax = (R)-INFINITY;
    R chk = 0;                       // fp32 mode: sticky non-finite detector (x * 0 accumulates NaN for inf / NaN)
    bool bad = false;
    int reg_days1 = 0, jump_days = 0;

    Slot slot[NS];
    SlotJ sj[JOURNAL ? NS : 1];
#pragma unroll(LIST != 0 || NS <= 8 ? NS : 1)
    for (int s = 0; s < NS; s++) { slot[s].W = (double)p.W0; slot[s].dd.init((double)p.W0); slot[s].aux = 0; slot[s].trades = 0; }
    if (JOURNAL) {
#pragma unroll(LIST != 0 || NS <= 8 ? NS : 1)
      for (int s = 0; s < NS; s++) { sj[s].pos[0] = sj[s].pos[1] = sj[s].pos[2] = 0; sj[s].vol = 0; sj[s].cnt = 0; }
    }
#pragma unroll(LIST != 0 || NS <= 8 ? NS : 1)
    for (int s = 0; s < NS; s++)
      if (TM_HAS(LETF_ST_PORTFOLIO) && s < p.n_strat && p.strat[s].type == LETF_ST_PORTFOLIO)
        for (int a = 0; a < kAssets; a++) slot[s].pv[a] = p.W0 * p.strat[s].w[a];
    Signals sg;
    sg.sma200.init(); sg.sma100.init(); sg.gain14.init(); sg.loss14.init(); sg.mom126.init(); sg.mean60.init();
    sg.v5.init(); sg.v10.init(); sg.v20.init(); sg.v60.init(); sg.dv20.init(); sg.sk60.init();
    for (int i = 0; i < 6; i++) sg.volhist[i] = nan_<R>();
    sg.P_prev = nan_<R>(); sg.vix_prev = nan_<R>(); sg.cP = sg.cS = sg.cA = 0;
    FastSig fs;
    for (int q = 0; q < 4; q++) { fs.a1[q] = 0; fs.a2[q] = 0; }
    fs.a3 = fs.d1 = fs.d2 = fs.p200 = fs.p100 = fs.g14 = fs.l14 = fs.m126 = (R)0;

    // ---- burn-in: generated days before the window (only the VIX recursion consumes them)
    for (int g = 0; g < so; g++) {
      int reg = regime_next(rg);
      R x[3] = {0, 0, 0}, vn = 0, in = 0;
      if (MARKET == kMarketBoot) {
        const X4 v = xs[g];
        boot_day(g, reg, v.x, v.y, v.z, shift, x, vn, in);
      } else if (two_pass && AUXSCR) {
        const X4 v = xs[g];
        x[0] = v.x + shift[0]; vn = v.w;
      } else if (two_pass) {
        market_day<false, true>(g, reg, js, x, vn, in);
        x[0] = xs[g].x + shift[0];
      } else {
        market_day<true, true>(g, reg, js, x, vn, in);
      }
      if (MARKET != LETF_MARKET_NORMAL) vix = vix_step(g, reg, vix, x[0], vn);
    }

    // ---- the simulated window
    // software prefetch: the scratch vector of day g + 1 and the ring entries this day retires are requested at the
    // top of the day, a whole day step ahead of their use (the loads are otherwise a fully exposed L2 round trip)
    X4 vnext; vnext.x = vnext.y = vnext.z = vnext.w = (R)0;
    R inext = (R)0;
    if (two_pass && T > 0) { vnext = xs[so]; if (AUXSCR) inext = is[so]; }
    // materialised-path mode: the day's series values are requested kSeriesAhead days ahead (a warp reads 256 B per
    // series and day; without the prefetch every day step is one exposed HBM round trip)
    constexpr int kSeriesAhead = LIST != 0 ? 2 : 4;      // (the typed 19-rule build has no registers for a deeper queue)
    double sq_er[MARKET == kMarketSeries ? kSeriesAhead : 1][kAssets];
    double sq_vix[MARKET == kMarketSeries ? kSeriesAhead : 1], sq_tlt[MARKET == kMarketSeries ? kSeriesAhead : 1];
    int sq_reg[MARKET == kMarketSeries ? kSeriesAhead : 1];
    R sq_cash[MARKET == kMarketSeries ? kSeriesAhead : 1];          // the frame's Cash_Ret column, when it has one
    auto series_fetch = [&](int slot, int day) {
      const int64_t o = (int64_t)day * ka.series.n_paths + col;
#pragma unroll
      for (int a = 0; a < kAssets; a++) sq_er[slot][a] = ka.series.asset_ret[a] ? ka.series.asset_ret[a][o] : 0.0;
      sq_vix[slot] = ka.series.vix ? ka.series.vix[o] : 15.0;
      sq_tlt[slot] = ka.series.tlt_ret ? ka.series.tlt_ret[o] : 0.0;
      sq_reg[slot] = ka.series.regime ? ka.series.regime[o] : 0;
      sq_cash[slot] = ka.series.cash_ret ? (R)ka.series.cash_ret[o] : (R)0;
    };
    if (MARKET == kMarketSeries) {
#pragma unroll
      for (int d = 0; d < kSeriesAhead; d++) if (d < T) series_fetch(d, d);
    }
    // record mode: everything a day reads from the producer's records - its NF fields and the values its rolling windows
    // retire (the same fields 5 ... 200 days back) - is staged through shared memory by asynchronous copies kStage - 1 days
    // ahead of its use (cp.async, one 4-byte element per thread and value: every thread owns its own slots, so no
    // barrier is needed and the copies cost neither registers nor scoreboard stalls).  Loads are coalesced: one column
    // per thread, every thread of a warp on the same day.  Slot v of day d: smem[((d % kStage) * kNV + v) * blockDim + tid].
    constexpr int kNV = kRecValues<NF>();
    // record field of the signal asset's return: the first one in the typed builds, else counted off the asset mask
    int fsig = 0;
    if (FSER && LIST == 0)
      for (int a = 0; a < kAssets; a++) if (a < p.sig_asset && ((AM >> a) & 1)) fsig++;
    const int64_t rnc = ka.rec_ncol, rfs = (int64_t)T * ka.rec_ncol;
    // the record element behind value v of day `day` (nullptr: no such value yet, the window is still filling)
    auto rec_src = [&](int v, int day) -> const float* {
      const float* r0 = ka.rec + (int64_t)day * rnc + col;
      if (v < NF) return r0 + (int64_t)v * rfs;
      const int lag = v == NF ? 5 : v == NF + 1 ? 10 : v == NF + 2 ? 20 : v == NF + 3 ? 60 : v == NF + 4 ? 126 :
                      v == NF + 5 ? 100 : v == NF + 6 ? 200 : v == NF + 7 ? 14 : 15;
      const int f = v <= NF + 3 ? fsig : (v == NF + 4 ? kFSpy : NF - 1);      // signal return | SPY return | SPY price
      return day >= lag ? r0 + (int64_t)f * rfs - (int64_t)lag * rnc : nullptr;
    };
    const bool rec_used[kNV > NF ? kNV - NF : 1] = {HAS_V5, HAS_V10, HAS_V20 || HAS_DV20, HAS_V60 || HAS_SKEW || HAS_MEAN60, NEED_RINGS,
                                                    HAS_SMA100, HAS_SMA200, HAS_RSI, HAS_RSI};
    R rec_cash_next = (CASHCOL && ka.rec_cash && T > 0) ? (R)ka.rec_cash[col] : (R)0;
#if defined(__CUDA_ARCH__)
    float* const stage = FSER ? letf_stage_smem() + threadIdx.x : nullptr;
    const int sblk = blockDim.x;
    // the copies of one day all read at (loop-invariant, warp-uniform) element offsets from that day's row of field 0:
    // one running pointer per thread, advanced by a row per day, instead of a 64-bit index product per value
    const float* issue_row = FSER ? ka.rec + col : nullptr;          // row of the next day to be issued
    int issue_slot = 0;                                              // its stage slot (day % kStage, kept incrementally)
    auto rec_off = [&](int v) -> int64_t {
      if (v < NF) return (int64_t)v * rfs;
      const int lag = v == NF ? 5 : v == NF + 1 ? 10 : v == NF + 2 ? 20 : v == NF + 3 ? 60 : v == NF + 4 ? 126 :
                      v == NF + 5 ? 100 : v == NF + 6 ? 200 : v == NF + 7 ? 14 : 15;
      const int f = v <= NF + 3 ? fsig : (v == NF + 4 ? kFSpy : NF - 1);
      return (int64_t)f * rfs - (int64_t)lag * rnc;
    };
    auto rec_lag = [](int v) -> int {
      return v < NF ? 0 : (v == NF ? 5 : v == NF + 1 ? 10 : v == NF + 2 ? 20 : v == NF + 3 ? 60 : v == NF + 4 ? 126 :
                           v == NF + 5 ? 100 : v == NF + 6 ? 200 : v == NF + 7 ? 14 : 15);
    };
    // all_valid: the caller knows that every window is full on that day (steady-state copy of the day body)
    auto stage_issue = [&](int day, bool all_valid) {
      if (day < T) {
        float* dst = stage + issue_slot * kNV * sblk;
#pragma unroll
        for (int v = 0; v < kNV; v++) {
          if (v >= NF && !rec_used[v - NF]) continue;
          if (all_valid || day >= rec_lag(v)) cp_async4(dst + v * sblk, issue_row + rec_off(v));
        }
        issue_row += rnc;
        issue_slot = issue_slot + 1 == kStage ? 0 : issue_slot + 1;
      }
      cp_async_commit();
    };
    int read_slot = 0;                                               // stage slot of the day being consumed
    auto stage_value = [&](int v, int) -> R { return (R)stage[(read_slot * kNV + v) * sblk]; };
    if (FSER) {
#pragma unroll
      for (int d = 0; d < kStage - 1; d++) stage_issue(d, false);
    }
#else
    auto stage_value = [&](int v, int day) -> R { const float* src = rec_src(v, day); return src ? (R)*src : (R)0; };
#endif
    // one simulated day.  steady_c = true_type: i >= kSteady, every rolling window is full (fp32 builds: the per-window
    // observation counts, their reciprocals and the warm-up predicates are compile-time constants in that copy of the body)
    constexpr int kSteady = 256;
    auto day_step = [&](const int i, auto steady_c) {
      constexpr bool STEADY = decltype(steady_c)::value;
      const int g = so + i;
      const int64_t sn = ka.series.n_paths;
      const X4 vcur = vnext;
      const R icur = inext;
      if (two_pass && i + 1 < T) { vnext = xs[g + 1]; if (AUXSCR) inext = is[g + 1]; }
      R pf_p200 = 0, pf_p100 = 0, pf_s126 = 0, pf_r60 = 0, pf_r20 = 0, pf_r10 = 0, pf_r5 = 0, pf_p14 = 0, pf_p15 = 0;
      if (FSER) {
#if defined(__CUDA_ARCH__)
        stage_issue(i + kStage - 1, STEADY);
        cp_async_wait<kStage - 1>();                    // day i's copies have landed (all but the newest kStage - 1 groups)
#endif
        // the values each window retires today (0 while the window is still filling)
        if (HAS_V5 && (STEADY || i >= 5)) pf_r5 = stage_value(NF, i);
        if (HAS_V10 && (STEADY || i >= 10)) pf_r10 = stage_value(NF + 1, i);
        if ((HAS_V20 || HAS_DV20) && (STEADY || i >= 20)) pf_r20 = stage_value(NF + 2, i);
        if ((HAS_V60 || HAS_SKEW || HAS_MEAN60) && (STEADY || i >= 60)) pf_r60 = stage_value(NF + 3, i);
        if (NEED_RINGS && (STEADY || i >= 126)) pf_s126 = stage_value(NF + 4, i);
        if (HAS_SMA100 && (STEADY || i >= 100)) pf_p100 = stage_value(NF + 5, i);
        if (HAS_SMA200 && (STEADY || i >= 200)) pf_p200 = stage_value(NF + 6, i);
        if (HAS_RSI && (STEADY || i >= 15)) { pf_p14 = stage_value(NF + 7, i); pf_p15 = stage_value(NF + 8, i); }
      } else if (!EXACT) {       // (the fp64 parity build keeps the loads at their use sites)
        if (NEED_RINGP) {
          const R* ring = ka.ringP + lane;
          if (HAS_SMA200 && i >= 200) pf_p200 = ring[roff(sg.cP)];
          if (HAS_SMA100 && i >= 100) { int c = sg.cP - 100; if (c < 0) c += kRingP; pf_p100 = ring[roff(c)]; }
          if (HAS_RSI && i >= 15) {
            int c1 = sg.cP - 14; if (c1 < 0) c1 += kRingP;
            int c2 = sg.cP - 15; if (c2 < 0) c2 += kRingP;
            pf_p14 = ring[roff(c1)]; pf_p15 = ring[roff(c2)];
          }
        }
        if (NEED_RINGS && i >= 126) pf_s126 = (ka.ringS + lane)[roff(sg.cS)];
        if (NEED_RINGA) {
          const R* ring = ka.ringA + lane;
          if (i >= 60 && (HAS_V60 || HAS_SKEW || HAS_MEAN60)) pf_r60 = ring[roff(sg.cA)];
          if (i >= 20 && (HAS_V20 || HAS_DV20)) { int c = sg.cA - 20; if (c < 0) c += kRingA; pf_r20 = ring[roff(c)]; }
          if (i >= 10 && HAS_V10) { int c = sg.cA - 10; if (c < 0) c += kRingA; pf_r10 = ring[roff(c)]; }
          if (i >= 5 && HAS_V5) { int c = sg.cA - 5; if (c < 0) c += kRingA; pf_r5 = ring[roff(c)]; }
        }
      }
      // series mode: take day i from the front of the prefetch queue, shift, request day i + kSeriesAhead
      R s_er[kAssets];
      int s_reg = 0;
      R s_cash = (R)0;
      if (MARKET == kMarketSeries) {
#pragma unroll
        for (int a = 0; a < kAssets; a++) s_er[a] = (R)sq_er[0][a];
        vix = (R)sq_vix[0]; s_reg = sq_reg[0]; s_cash = sq_cash[0];
      }
      R rec_pspy = (R)100;
      if (FSER) {
        int f = 0;
#pragma unroll
        for (int a = 0; a < kAssets; a++) s_er[a] = ((AM >> a) & 1) ? stage_value(f++, i) : (R)0;
        const R v = stage_value(NF - 2, i);
        vix = M::abs_(v); s_reg = v < (R)0 ? 1 : 0;              // VIX >= 10: the sign bit carries the regime
        rec_pspy = stage_value(NF - 1, i);
#if defined(__CUDA_ARCH__)
        read_slot = read_slot + 1 == kStage ? 0 : read_slot + 1;
#endif
      }
      const int reg = SERIES ? s_reg : regime_next(rg);
      R x[3] = {0, 0, 0}, vn = 0, in = 0;
      if (FSER) {
      } else if (MARKET == kMarketSeries) {
        x[2] = (R)sq_tlt[0];
#pragma unroll
        for (int d = 0; d + 1 < kSeriesAhead; d++) {
#pragma unroll
          for (int a = 0; a < kAssets; a++) sq_er[d][a] = sq_er[d + 1][a];
          sq_vix[d] = sq_vix[d + 1]; sq_tlt[d] = sq_tlt[d + 1]; sq_reg[d] = sq_reg[d + 1]; sq_cash[d] = sq_cash[d + 1];
        }
        if (i + kSeriesAhead < T) series_fetch(kSeriesAhead - 1, i + kSeriesAhead);
      } else if (MARKET == kMarketBoot) {
        boot_day(g, reg, vcur.x, vcur.y, vcur.z, shift, x, vn, in);
      } else if (two_pass) {
        if (AUXSCR) { vn = vcur.w; in = icur; }
        else market_day<false, true>(g, reg, js, x, vn, in);
        x[0] = vcur.x + shift[0]; x[1] = vcur.y + shift[1]; x[2] = vcur.z + shift[2];
      } else {
        market_day<true, true>(g, reg, js, x, vn, in);
      }
      R rf = p.const_rf;
      if (MARKET != LETF_MARKET_NORMAL && !SERIES) {
        vix = vix_step(g, reg, vix, x[0], vn);
        R irx = clipx(p.irx_base[reg] + ((R)0 + (R)0.5 * in), (R)0, (R)15);   // engine.py:117-123
        rf = EXACT ? clipx(irx, (R)0, (R)20) / (R)100 : clipx(irx, (R)0, (R)20) * (R)0.01;                                // engine.py:516
      }
      R spy = x[0], qqq = x[1], tlt = x[2];
      if (TAPEIO && ka.log_market && !ka.trial_ids && ka.log_trial == k) {
        // the generated underlying returns, before the jump overlay (what calibration.py:480-590 returns)
        for (int a = 0; a < 3; a++) ka.log_market[(int64_t)i * 3 + a] = (double)x[a];
      }
      if (EXACT) {
        if (!isfinite_(spy) || !isfinite_(qqq) || !isfinite_(tlt)) flags |= LETF_FLAG_NONFINITE_RET;
      } else {
        chk = M::fma_(spy, (R)0, chk); chk = M::fma_(qqq, (R)0, chk); chk = M::fma_(tlt, (R)0, chk);
      }
      if (MARKET != LETF_MARKET_NORMAL) {             // constant VIX / regime 0 in the iid-normal shapes: checked once below
        if (!isfinite_(vix) || vix <= (R)0) flags |= LETF_FLAG_BAD_VIX;
        reg_days1 += reg;
        vmin = rmin(vmin, vix); vmax = rmax(vmax, vix);
      }

      // ---- latent stress + jump overlay (calibration.py:787-830, engine.py:502-512)
      if (!SERIES && MARKET != LETF_MARKET_NORMAL && p.has_stress) {   // params.py: no stress model with the iid-normal shapes
        const DevStress<R>& sp = p.stress[reg];
        R jump = 0, sign = 1;
        if (i == 0) {
          liq = sp.liq_mu; cred = sp.cr_mu;      // no draws on the first day (calibration.py:798-803)
          if (rec) { tp_put(ka.tape.liq_n, 0, 0); tp_put(ka.tape.cred_n, 0, 0); tp_put(ka.tape.jump_u, 0, 1); tp_put(ka.tape.jump_t, 0, 0); tp_put(ka.tape.sign, 0, 1); }
        } else {
          R ln, cn, ju;
          U4 b; b.x = b.y = b.z = b.w = 0;
          if (use_tape) { ln = tp_get(ka.tape.liq_n, i); cn = tp_get(ka.tape.cred_n, i); ju = tp_get(ka.tape.jump_u, i); sign = tp_get(ka.tape.sign, i); }
          else {
            b = rng_block(key, (uint32_t)g, STREAM_STRESS);
            M::box_muller(b.x, b.y, ln, cn);
            ju = (R)(b.z >> 8) * (R)(1.0 / 16777216.0);    // U[0,1) on a 24-bit grid, like numpy's float32 path
            sign = (b.w & 1u) ? (R)1 : (R)-1;
            if (rec) { tp_put(ka.tape.liq_n, i, ln); tp_put(ka.tape.cred_n, i, cn); tp_put(ka.tape.jump_u, i, ju); tp_put(ka.tape.sign, i, sign); }
          }
          liq = clipx(sp.liq_mu + sp.liq_phi * (liq - sp.liq_mu) + ((R)0 + sp.liq_sigma * ln), (R)0, (R)3);
          cred = clipx(sp.cr_mu + sp.cr_phi * (cred - sp.cr_mu) + ((R)0 + sp.cr_sigma * cn), (R)0, (R)3);
          R vamp = rmax(EXACT ? (vix - (R)25) / (R)30 : (vix - (R)25) * (R)(1.0 / 30.0), (R)0);
          R jp = clipx(sp.jbase + (R)0.008 * vamp + (R)0.003 * liq, (R)0, (R)0.01);
          if (ju < jp || rec) {
            // record mode draws the jump size every day: antithetic twins share a tape column but
            // may jump on different days, so the column must hold the variate either twin would use
            R t5;
            if (use_tape) t5 = tp_get(ka.tape.jump_t, i);
            else {
              U4 bj = rng_block(key, (uint32_t)g, STREAM_JUMP);
              t5 = student_t_block<R, KEYREF>(bj, (R)5, 5, key, (uint32_t)g, STREAM_JUMP);
              if (rec) tp_put(ka.tape.jump_t, i, t5);
            }
            if (ju < jp) {
              jump = M::abs_(t5) * sp.jscale;
              if (jump > (R)0) jump_days++;
            }
          }
        }
        R sjv = jump * sign;
        spy = clipr(spy - sjv, (R)-0.95, (R)3.0);
        qqq = clipr(qqq - (R)1.12 * sjv, (R)-0.95, (R)4.0);
      }

      // ---- financing spread shared by all assets (engine.py:328-342, calibration.py:833-857)
      R spread_common = 0;
      if (!SERIES && p.has_funding) {
        R irx_level = rf * (R)100;
        R term = p.has_stress ? (R)1.25 - (R)1.1 * clipx(cred, (R)0, (R)3) : (vix > (R)30 ? (R)-0.25 : (R)1.10);
        R tnx = irx_level + term;
        R sv = rmax(vix - (R)20, (R)0);
        R inv = rmax(-(tnx - irx_level), (R)0);
        R spd = p.f_base + p.f_vix * sv + p.f_inv * inv;
        if (p.has_stress) spd += p.f_liq * clipx(liq, (R)0, (R)3) + p.f_cr * clipx(cred, (R)0, (R)3);
        spread_common = clipx(spd, p.f_min, p.f_max);
      }
      const R vmult = clipx(M::sqrt_(EXACT ? vix / (R)20 : vix * (R)0.05), (R)0.5, (R)1.5);
      const R rmult = reg == 0 ? (R)1 : (R)1.15;
      const R vrmult = vmult * rmult;

      // ---- the six ETF layers (engine.py:360-420)
      R ret[kAssets];
#pragma unroll
      for (int a = 0; a < kAssets; a++) ret[a] = 0;
#pragma unroll
      for (int a = 0; a < NA; a++) {             // a = slot; aid = asset id (identical when NA == 6)
        const int aid = NA == kAssets ? a : p.act_id[a];
        R er;
        if (SERIES) {
          if (FSER ? !((AM >> a) & 1) : !ka.series.asset_ret[a]) continue;
          er = s_er[a];
        } else {
          if (AM != 0 && NA == kAssets ? !((AM >> a) & 1) : !p.act_on[a]) continue;
          const DevAsset<R>& ap = p.act[a];
          R ub = ap.und == 1 ? qqq : (ap.und == 2 ? tlt : spy);
          R u = EXACT ? (ap.beta_applies ? ub * ap.beta : ub) : ub * ap.beta_eff;
          R spread = p.has_funding ? spread_common : ap.borrow_spread;
          R lev;
          if (!EXACT && MARKET == LETF_MARKET_NORMAL && !p.has_funding) {
            lev = ap.lev * u - ap.carry_const;            // constant rf, no funding model: the carry is invariant
          } else if (EXACT) {
            R borrow = 0;
            if (ap.bl > (R)0) borrow = ap.bl * rmax(rf + spread, (R)0) / (R)252;
            lev = ap.lev * u - ap.er_daily - borrow;
          } else {
            lev = ap.lev * u - (ap.er_daily + ap.bl_252 * rmax(rf + spread, (R)0));
          }
          R te_raw = 0;
          if (TAPEIO ? i >= 1 : true) {      // Philox-only builds draw on day 0 too and discard it below (no branch)
            R innov;
            if (use_tape) innov = tp_get(ka.tape.te_t, (int64_t)aid * T + i);
            else {
              U4 b = rng_block(key, (uint32_t)g, STREAM_TE0 + aid);
              // builds with a compile-time asset mask also fix the tracking-error dof at cfg.ASSETS' defaults
              // (t(5) for the four LETFs, t(10) for SPY / QQQ; letf_api.cu checks it): no sampler dispatch
              // (bit 6 of AM = "tracking-error dof at run time": the typed builds for calibrated residual models)
              if (AM != 0 && !(AM & 0x40) && NA == kAssets && !EXACT) innov = a >= 4 ? student_t10<R>(b) : student_t5<R>(b);
              else innov = student_t_block<R, KEYREF>(b, ap.te_df, ap.te_fast_k, key, (uint32_t)g, STREAM_TE0 + aid);
              if (rec) tp_put(ka.tape.te_t, (int64_t)aid * T + i, innov);
            }
            R ds = u < (R)0 ? ap.te_down : (R)0.98;
            R mm = ((R)1 + (R)0.3 * M::abs_(u)) * ds;
            if (EXACT) {
              innov *= ap.te_base * vmult * rmult * (R)1;
              te_raw = ap.te_rho * te_prev[a] + ap.te_arw * innov * mm;
            } else {
              te_raw = ap.te_rho * te_prev[a] + (innov * vrmult) * (ap.te_scale * mm);
              if (!TAPEIO) te_raw = i >= 1 ? te_raw : (R)0;
            }
          } else if (rec) tp_put(ka.tape.te_t, (int64_t)aid * T, 0);
          te_prev[a] = te_raw;
          const R alpha = (R)(2.0 / 127.0);
          bias[a] = alpha * te_raw + ((R)1 - alpha) * bias[a];
          R te = te_raw - bias[a];
          if (ap.te_clip > (R)0) te = ap.te_clip * M::tanh_(M::div_(te, ap.te_clip));
          er = EXACT ? clipr(lev + te, (R)-0.999, (R)10) : clipfast(lev + te, (R)-0.999, (R)10);
        }
        if (NA == kAssets) ret[a] = er;
        if (FSER) continue;                     // prices and buy&hold statistics belong to the producer
        if (EXACT) {
          cp[a] = i == 0 ? (1.0 + (double)er) : cp[a] * (1.0 + (double)er);
          double pr = cp[a] * 100.0;
          if (!isfinite_(er) || !isfinite_(pr) || !(pr > 0.0)) bad = true;
        } else {
          cp[a] = ::fma(cp[a], (double)er, cp[a]);      // a NaN return or a product that reached 0 stays there
        }
        if (p.act_bench[a]) {                   // strategy.py:85-95 buy & hold: W = W0 * cumprod(1 + r)
          if (EXACT) {
            if (i == 0) ddA[a].init((double)p.W0 * cp[a]);
            else ddA[a].step((double)p.W0 * cp[a], (R)1 + er);
          } else ddA[a].step_ret(er);                   // starts "open" (q = inf): day 0 lands on q = 1
        }
      }
      {
        if (EXACT) {
          cp_tlt = i == 0 ? (1.0 + (double)tlt) : cp_tlt * (1.0 + (double)tlt);
          if (!isfinite_(tlt) || !(cp_tlt > 0.0)) bad = true;
        } else if (!FSER && p.bench_tlt) cp_tlt = ::fma(cp_tlt, (double)tlt, cp_tlt);
        if (!FSER && p.bench_tlt) {
          if (EXACT) {
            if (i == 0) ddT.init((double)p.W0 * cp_tlt);
            else ddT.step((double)p.W0 * cp_tlt, (R)1 + tlt);
          } else ddT.step_ret(tlt);
        }
      }
      R cash = (MARKET == kMarketSeries && ka.series.cash_ret) ? s_cash : p.cash_daily[reg];
      if (CASHCOL && ka.rec_cash) {       // requested a day ahead (one plain load per day: the column is not worth a staging slot)
        cash = rec_cash_next;
        if (i + 1 < T) rec_cash_next = (R)ka.rec_cash[(int64_t)(i + 1) * ka.rec_ncol + col];
      }
      const R Pspy = FSER ? rec_pspy : (NA == kAssets ? (R)(cp[NA == kAssets ? 4 : 0] * 100.0) : (R)100);

      // ---- signals as of yesterday, then roll the windows forward with today's values
      R s_sma200 = nan_<R>(), s_sma200f = nan_<R>(), s_sma100 = nan_<R>(), s_rsi = nan_<R>(), s_mom = nan_<R>();
      R s_v5 = nan_<R>(), s_v10 = nan_<R>(), s_v20 = nan_<R>(), s_v20f = nan_<R>(), s_v60 = nan_<R>();
      R s_dv20 = nan_<R>(), s_dv20f = nan_<R>(), s_sk = nan_<R>(), s_mean60 = nan_<R>(), s_vm = nan_<R>(), s_v20_today = nan_<R>();
      const R P_prev = sg.P_prev, vix_prev = sg.vix_prev;
      const R ra = LIST != 0 ? ret[0] : pick6(ret, p.sig_asset);
      if (EXACT) {
        if (i >= 1) {
          if (HAS_SMA200) { s_sma200 = sg.sma200.mean(1); s_sma200f = sg.sma200.mean(200); }
          if (HAS_SMA100) s_sma100 = sg.sma100.mean(1);
          if (HAS_RSI) {
            R gm = sg.gain14.mean(14), lm = sg.loss14.mean(14);
            R rs = M::div_(gm, lm);
            s_rsi = (R)100 - M::div_((R)100, (R)1 + rs);
          }
          if (HAS_MOM) s_mom = sg.mom126.sum(1);
          if (HAS_V5) s_v5 = sg.v5.stdev(1) * V;
          if (HAS_V10) s_v10 = sg.v10.stdev(1) * V;
          if (HAS_V20) { s_v20 = sg.v20.stdev(1) * V; s_v20f = sg.v20.stdev(20) * V; }
          if (HAS_V60) s_v60 = sg.v60.stdev(1) * V;
          if (HAS_DV20) { s_dv20 = sg.dv20.stdev(1) * V; s_dv20f = sg.dv20.stdev(20) * V; }
          if (HAS_SKEW) s_sk = sg.sk60.skew(1);
          if (HAS_MEAN60) s_mean60 = sg.mean60.mean(1) * (R)252;
        }
        if (HAS_VOLMOM) {
          // vol_slow.pct_change(5): shifted 20-day vol today vs five rows earlier
          for (int q = 5; q >= 1; q--) sg.volhist[q] = sg.volhist[q - 1];
          sg.volhist[0] = s_v20;
          s_vm = M::div_(sg.volhist[0], sg.volhist[5]) - (R)1;
        }
        // roll forward (pandas: removals first, then the new observation)
        if (NEED_RINGP) {
          R* ring = ka.ringP + lane;
          if (HAS_SMA200) { if (i >= 200) sg.sma200.remove(EXACT ? ring[roff(sg.cP)] : pf_p200); sg.sma200.add(Pspy); }
          if (HAS_SMA100) { if (i >= 100) { int c = sg.cP - 100; if (c < 0) c += kRingP; sg.sma100.remove(EXACT ? ring[roff(c)] : pf_p100); } sg.sma100.add(Pspy); }
          if (HAS_RSI) {
            if (i >= 14) {
              R dg = 0, dl = -(R)0;
              if (i >= 15) {
                int c1 = sg.cP - 14; if (c1 < 0) c1 += kRingP;
                int c2 = sg.cP - 15; if (c2 < 0) c2 += kRingP;
                R dlt = ring[roff(c1)] - ring[roff(c2)];
                dg = dlt > (R)0 ? dlt : (R)0;
                dl = -(dlt < (R)0 ? dlt : (R)0);
              }
              sg.gain14.remove(dg); sg.loss14.remove(dl);
            }
            R dlt = i == 0 ? nan_<R>() : Pspy - P_prev;
            sg.gain14.add(dlt > (R)0 ? dlt : (R)0);
            sg.loss14.add(-(dlt < (R)0 ? dlt : (R)0));
          }
          ring[roff(sg.cP)] = Pspy;
          sg.cP = sg.cP + 1 == kRingP ? 0 : sg.cP + 1;
        }
        if (NEED_RINGS) {
          R* ring = ka.ringS + lane;
          if (i >= 126) sg.mom126.remove(EXACT ? ring[roff(sg.cS)] : pf_s126);
          sg.mom126.add(ret[4]);
          ring[roff(sg.cS)] = ret[4];
          sg.cS = sg.cS + 1 == kRingS ? 0 : sg.cS + 1;
        }
        if (NEED_RINGA) {
          R* ring = ka.ringA + lane;
          R r60 = pf_r60, r20 = pf_r20, r10 = pf_r10, r5 = pf_r5;
          if (EXACT) {
            if (i >= 60 && (HAS_V60 || HAS_SKEW || HAS_MEAN60)) r60 = ring[roff(sg.cA)];
            if (i >= 20 && (HAS_V20 || HAS_DV20)) { int c = sg.cA - 20; if (c < 0) c += kRingA; r20 = ring[roff(c)]; }
            if (i >= 10 && HAS_V10) { int c = sg.cA - 10; if (c < 0) c += kRingA; r10 = ring[roff(c)]; }
            if (i >= 5 && HAS_V5) { int c = sg.cA - 5; if (c < 0) c += kRingA; r5 = ring[roff(c)]; }
          }
          if (HAS_V5) { if (i >= 5 && sg.v5.remove(r5)) revar_(sg.v5, ring, nth, sg.cA, 5, false); sg.v5.add(ra); }
          if (HAS_V10) { if (i >= 10 && sg.v10.remove(r10)) revar_(sg.v10, ring, nth, sg.cA, 10, false); sg.v10.add(ra); }
          if (HAS_V20) { if (i >= 20 && sg.v20.remove(r20)) revar_(sg.v20, ring, nth, sg.cA, 20, false); sg.v20.add(ra); }
          if (HAS_V60) { if (i >= 60 && sg.v60.remove(r60)) revar_(sg.v60, ring, nth, sg.cA, 60, false); sg.v60.add(ra); }
          if (HAS_DV20) { if (i >= 20 && sg.dv20.remove(r20 < (R)0 ? r20 : (R)0)) revar_(sg.dv20, ring, nth, sg.cA, 20, true); sg.dv20.add(ra < (R)0 ? ra : (R)0); }
          if (HAS_SKEW) { if (i >= 60) sg.sk60.remove(r60); sg.sk60.add(ra); }
          if (HAS_MEAN60) { if (i >= 60) sg.mean60.remove(r60); sg.mean60.add(ra); }
          ring[roff(sg.cA)] = ra;
          sg.cA = sg.cA + 1 == kRingA ? 0 : sg.cA + 1;
          if (TM_HAS(LETF_ST_VOL_TARGETING)) s_v20_today = sg.v20.stdev(20) * V;   // strategy.py:255 (unshifted)
        }
      } else {
        // ---- fp32: sliding power sums (FastSig).  n(w) = min(i, w) observations are in the state "as of yesterday"
        R ri = (R)0, rim1 = (R)0;
        if (!STEADY) { ri = i > 0 ? M::div_((R)1, (R)i) : (R)0; rim1 = i > 1 ? M::div_((R)1, (R)(i - 1)) : (R)0; }
        auto inv = [&](int w) -> R { return (STEADY || i >= w) ? (R)(1.0 / w) : ri; };
        auto invm1 = [&](int w) -> R { return (STEADY || i >= w) ? (R)(1.0 / (w - 1)) : rim1; };
        auto sdev = [&](R s1, R s2, int w) -> R {      // rolling(w).std() * sqrt(252) from the window's power sums
          const R var = rmax(M::fma_(-s1 * inv(w), s1, s2), (R)0) * invm1(w);
          return M::sqrt_(var) * V;
        };
        if (STEADY || i >= 1) {
          if (HAS_SMA200) { s_sma200 = fs.p200 * inv(200); if (STEADY || i >= 200) s_sma200f = s_sma200; }
          if (HAS_SMA100) s_sma100 = fs.p100 * inv(100);
          if (HAS_RSI && (STEADY || i >= 14)) s_rsi = (R)100 - M::div_((R)100, (R)1 + M::div_(fs.g14, fs.l14));
          if (HAS_MOM) s_mom = fs.m126;
          if (HAS_MEAN60) s_mean60 = fs.a1[3] * inv(60) * (R)252;
        }
        if (STEADY || i >= 2) {
          if (HAS_V5) s_v5 = sdev(fs.a1[0], fs.a2[0], 5);
          if (HAS_V10) s_v10 = sdev(fs.a1[1], fs.a2[1], 10);
          if (HAS_V20) { s_v20 = sdev(fs.a1[2], fs.a2[2], 20); if (STEADY || i >= 20) s_v20f = s_v20; }
          if (HAS_V60) s_v60 = sdev(fs.a1[3], fs.a2[3], 60);
          if (HAS_DV20) { s_dv20 = sdev(fs.d1, fs.d2, 20); if (STEADY || i >= 20) s_dv20f = s_dv20; }
        }
        if (HAS_SKEW && (STEADY || i >= 3)) {
          // bias-corrected sample skewness (pandas rolling.skew) of the last n = min(i, 60) returns
          const R n = (STEADY || i >= 60) ? (R)60 : (R)i, in_ = inv(60);
          const R k = (STEADY || i >= 60) ? (R)1.0258258469409323                      // sqrt(n (n-1)) / (n-2) at n = 60
                                          : M::div_(M::sqrt_(n * (n - (R)1)), n - (R)2);
          const R m = fs.a1[3] * in_;
          const R M2 = M::fma_(-fs.a1[3], m, fs.a2[3]);                                // sum (r - m)^2
          const R M3 = M::fma_((R)2 * m * m, fs.a1[3], M::fma_((R)-3 * m, fs.a2[3], fs.a3));   // sum (r - m)^3
          const R B = M2 * in_;
          const R rb = M::rsqrt_(rmax(B, (R)1e-30));
          s_sk = B <= (R)1e-14 ? nan_<R>() : k * (M3 * in_) * (rb * rb * rb);
        }
        if (HAS_VOLMOM) {
          for (int q = 5; q >= 1; q--) sg.volhist[q] = sg.volhist[q - 1];
          sg.volhist[0] = s_v20;
          s_vm = M::div_(sg.volhist[0], sg.volhist[5]) - (R)1;
        }
        // roll forward: subtract the value each window retires (0 while the window is still filling), add today's
        if (NEED_RINGP) {
          if (HAS_SMA200) fs.p200 += Pspy - pf_p200;
          if (HAS_SMA100) fs.p100 += Pspy - pf_p100;
          if (HAS_RSI) {
            const R dold = pf_p14 - pf_p15;                       // price difference of day i - 14 (0 before day 15)
            const R dnew = (STEADY || i >= 1) ? Pspy - P_prev : (R)0;
            fs.g14 += rmax(dnew, (R)0) - rmax(dold, (R)0);
            fs.l14 += rmax(-dnew, (R)0) - rmax(-dold, (R)0);
          }
          if (!FSER) {
            R* ring = ka.ringP + lane;
            ring[roff(sg.cP)] = Pspy;
            sg.cP = sg.cP + 1 == kRingP ? 0 : sg.cP + 1;
          }
        }
        if (NEED_RINGS) {
          fs.m126 += ret[4] - pf_s126;
          if (!FSER) {
            R* ring = ka.ringS + lane;
            ring[roff(sg.cS)] = ret[4];
            sg.cS = sg.cS + 1 == kRingS ? 0 : sg.cS + 1;
          }
        }
        if (NEED_RINGA) {
          const R olds[4] = {pf_r5, pf_r10, pf_r20, pf_r60};
          const bool need[4] = {HAS_V5, HAS_V10, HAS_V20, HAS_V60 || HAS_SKEW || HAS_MEAN60};
#pragma unroll
          for (int q = 0; q < 4; q++) if (need[q]) {
            fs.a1[q] += ra - olds[q];
            fs.a2[q] = M::fma_(-olds[q], olds[q], M::fma_(ra, ra, fs.a2[q]));
          }
          if (HAS_SKEW) fs.a3 = M::fma_(-pf_r60 * pf_r60, pf_r60, M::fma_(ra * ra, ra, fs.a3));
          if (HAS_DV20) {
            const R dn = rmin(ra, (R)0), dol = rmin(pf_r20, (R)0);
            fs.d1 += dn - dol;
            fs.d2 = M::fma_(-dol, dol, M::fma_(dn, dn, fs.d2));
          }
          if (!FSER) {
            R* ring = ka.ringA + lane;
            ring[roff(sg.cA)] = ra;
            sg.cA = sg.cA + 1 == kRingA ? 0 : sg.cA + 1;
          }
          if (TM_HAS(LETF_ST_VOL_TARGETING) && (STEADY || i >= 19)) {                  // strategy.py:255 (unshifted): n = min(i + 1, 20)
            const R var = rmax(M::fma_(-fs.a1[2] * (R)(1.0 / 20), fs.a1[2], fs.a2[2]), (R)0) * (R)(1.0 / 19);
            s_v20_today = M::sqrt_(var) * V;
          }
        }
        // drift control: every kRefresh days rebuild the sums from the windows' history (today's value included)
        if (STEADY && (i & (kRefresh - 1)) == kRefresh - 1) {
          auto hist = [&](int which, int back) -> R {          // value `back` days ago (0 = today) of series 0: signal return, 1: SPY return, 2: SPY price
            if (FSER) {
              const int f = which == 0 ? fsig : (which == 1 ? kFSpy : NF - 1);
              return (R)ka.rec[((int64_t)f * T + (i - back)) * ka.rec_ncol + col];
            }
            const int depth = which == 0 ? kRingA : (which == 1 ? kRingS : kRingP);
            int c = (which == 0 ? sg.cA : (which == 1 ? sg.cS : sg.cP)) - 1 - back;    // cursors already point at tomorrow's slot
            if (c < 0) c += depth;
            const R* ring = which == 0 ? ka.ringA : (which == 1 ? ka.ringS : ka.ringP);
            return (ring + lane)[roff(c)];
          };
          if (NEED_RINGA) {
            R t1 = 0, t2 = 0, t3 = 0, e1 = 0, e2 = 0;
            for (int b = 0; b < 60; b++) {
              const R v = hist(0, b);
              t1 += v; t2 = M::fma_(v, v, t2); t3 = M::fma_(v * v, v, t3);
              if (b < 20) { const R dn = rmin(v, (R)0); e1 += dn; e2 = M::fma_(dn, dn, e2); }
              if (b == 4) { fs.a1[0] = t1; fs.a2[0] = t2; }
              if (b == 9) { fs.a1[1] = t1; fs.a2[1] = t2; }
              if (b == 19) { fs.a1[2] = t1; fs.a2[2] = t2; fs.d1 = e1; fs.d2 = e2; }
            }
            fs.a1[3] = t1; fs.a2[3] = t2; fs.a3 = t3;
          }
          if (NEED_RINGS) { R t = 0; for (int b = 0; b < 126; b++) t += hist(1, b); fs.m126 = t; }
          if (NEED_RINGP) {
            R t = 0, gq = 0, lq = 0, newer = 0;
            for (int b = 0; b < 200; b++) {
              const R v = hist(2, b);
              if (HAS_RSI && b >= 1 && b <= 14) { const R d = newer - v; gq += rmax(d, (R)0); lq += rmax(-d, (R)0); }
              newer = v;
              t += v;
              if (b == 99) fs.p100 = t;
            }
            fs.p200 = t;
            if (HAS_RSI) { fs.g14 = gq; fs.l14 = lq; }
          }
        }
      }
      sg.P_prev = Pspy; sg.vix_prev = vix;

      // ---- strategy rules (strategy.py:55-1277) + running statistics (engine.py:578-590)
      // (buy & hold needs no slot work: its equity is W0 * cp[asset], tracked above)
      auto slot_body = [&](const int s, auto tyc_, auto asc_) {
        constexpr int TYC = decltype(tyc_)::value;      // this slot's compile-time rule (typed lists), -1 = dynamic, -2 = empty
        constexpr int ASC = decltype(asc_)::value;      // its compile-time traded asset, -1 = read p.strat[s].asset
        if (TYC == -2 || TYC == LETF_ST_BENCHMARK || TYC == LETF_ST_BENCHMARK_TLT) return;
        if (TYC < 0 && s >= p.n_strat) return;
        const DevStrategy<R>& st = p.strat[s];
        if (ROUTED && st.route < 0) return;
        Slot& sl = slot[s];
        const int ty = TYC >= 0 ? TYC : st.type;
        const R r_asset = ASC >= 0 ? ret[ASC >= 0 ? ASC : 0] : pick6(ret, st.asset);     // return of the traded asset
        const R lev_asset = ASC >= 0 ? p.assets[ASC >= 0 ? ASC : 0].lev : p.assets[st.asset].lev;
        R growth = (R)1;          // W_i / W_{i-1}
        if (ty == LETF_ST_BENCHMARK || ty == LETF_ST_BENCHMARK_TLT) return;
        R r_strat = 0;
        if (TM_HAS(LETF_ST_SMA) || TM_HAS(LETF_ST_SMA_BAND)) if (ty == LETF_ST_SMA || ty == LETF_ST_SMA_BAND) {
          R thr = ty == LETF_ST_SMA ? s_sma200 : s_sma200 * ((R)1 - st.p[1]);
          bool buy = P_prev >= thr, sell = P_prev < thr;
          int pos_prev = (int)sl.aux, pos = pos_prev == 0 ? (buy ? 1 : 0) : (sell ? 0 : 1);
          r_strat = pos == 1 ? r_asset : cash;
          if (EXACT) {
            if (pos != pos_prev) {
              sl.trades++;
              if (p.apply_costs) r_strat -= p.spread_cost[reg] + p.rebalance_cost * lev_asset * M::abs_(r_asset) + (R)0;
            }
          } else {
            const bool sw = pos != pos_prev;
            sl.trades += sw ? 1 : 0;
            const R tc = p.spread_cost[reg] + p.rebalance_cost * lev_asset * M::abs_(r_asset);
            if (p.apply_costs) r_strat -= sw ? tc : (R)0;
          }
          sl.aux = (R)pos;
        }
        if (TM_HAS(LETF_ST_PORTFOLIO)) if (ty == LETF_ST_PORTFOLIO) {
          R* pvv = sl.pv;
          R total = 0;
#define LETF_PORT_ASSET(a) ((AM == 0 || ((AM >> (a)) & 1)) && st.w[a] != (R)0)
          for (int a = 0; a < kAssets; a++) if (LETF_PORT_ASSET(a)) { pvv[a] = pvv[a] * ((R)1 + ret[a]); total += pvv[a]; }
          double w_new = (double)total;
          // i % rebalance_freq == 0 (strategy.py:182) as a countdown kept in the slot's spare field
          sl.aux += (R)1;
          const bool rebalance = sl.aux >= st.p[0];
          if (rebalance) sl.aux = (R)0;
          if (rebalance) {
            R wt = 0, lt = 0;
            for (int a = 0; a < kAssets; a++) if (LETF_PORT_ASSET(a)) {
              R cw = M::div_(pvv[a], total);
              wt += M::abs_(cw - st.w[a]);
              R tl = p.assets[a].lev, emb = (R)1;
              if (tl > (R)1) {
                R uu = M::div_(ret[a], tl);
                emb = M::abs_((R)1 + tl * uu) > (R)0.01 ? M::div_(tl * ((R)1 + uu), (R)1 + tl * uu) : tl;
              }
              lt += M::div_(M::abs_(emb - tl), tl) * cw;
            }
            R turn = wt + lt;
            if (p.apply_costs && turn > (R)0.01) { total -= turn * p.rebalance_cost * total; w_new = (double)total; }
            for (int a = 0; a < kAssets; a++) if (LETF_PORT_ASSET(a)) { pvv[a] = total * st.w[a]; sl.trades++; }
          }
#undef LETF_PORT_ASSET
          growth = EXACT ? (R)(w_new / sl.W) : M::div_((R)w_new, (R)sl.W);
          sl.W = w_new;
          sl.dd.step(sl.W, growth);
          return;
        }
        if (TM_HAS(LETF_ST_COMPOSITE)) if (ty == LETF_ST_COMPOSITE) {
          int score = (P_prev > s_sma200f ? 1 : 0) + (((R)40 < s_rsi && s_rsi < (R)80) ? 1 : 0) + (vix_prev < st.p[2] ? 1 : 0);
          int target = score == 3 ? 2 : (score == 2 ? 1 : 0);
          r_strat = target == 2 ? r_asset : (target == 1 ? (ASC >= 0 ? ret[4] : pick6(ret, st.asset2)) : cash);
          if (target != (int)sl.aux) {
            sl.trades++;
            if (JOURNAL && st.taxable) {
              int ta = target == 2 ? st.asset : (target == 1 ? st.asset2 : 4);
              // journal keyed by asset NAME in the reference: legs that alias the same ticker share a position
              int which = ta == st.asset ? 0 : (ta == st.asset2 ? 1 : 2);
              jlog(sj[s], which, s, i, ta, target == 0 ? (R)0 : (R)1, sl.W, (R)(cp[ta] * 100.0));
            }
            sl.aux = (R)target;
          }
        }
        if (HAS_VOL) if ((kVolTypes >> ty) & 1u) {
          R alloc = 0, thresh = (R)0.0001;
          bool sticky = false, cashleg = true, skip = false;
          switch (ty) {
            case LETF_ST_VOL_TARGETING: if (TM_HAS(LETF_ST_VOL_TARGETING)) {
              R cv = s_v20_today;
              alloc = (isnan_(cv) || cv < (R)0.01) ? (R)1 : clipr(M::div_(st.p[0], cv), (R)0.2, (R)2.0);
              cashleg = false;
            } break;
            case LETF_ST_ADAPTIVE_VOL: if (TM_HAS(LETF_ST_ADAPTIVE_VOL)) {
              R cv = s_v20f, cs = s_sma200f;
              if (isnan_(cv) || isnan_(cs) || cv < (R)0.001) skip = true;
              else alloc = clipr(M::div_(P_prev > cs ? st.p[0] : st.p[1], cv), (R)0, (R)1);
            } break;
            case LETF_ST_DOWNSIDE_VOL: if (TM_HAS(LETF_ST_DOWNSIDE_VOL)) {
              R dv = s_dv20f;
              alloc = (isnan_(dv) || dv < (R)0.001) ? (R)1 : clipr(M::div_(st.p[0], dv), (R)0, (R)1.5);
            } break;
            case LETF_ST_CONVEX_VOL: if (TM_HAS(LETF_ST_CONVEX_VOL)) {
              R v = s_v20; if (isnan_(v) || v < (R)0.001) v = (R)0.5;
              R al = M::div_(st.p[0], v);
              if (P_prev > s_sma200) al = M::pow_(al, st.p[1]);
              alloc = clipr(al, (R)0, (R)1);
            } break;
            case LETF_ST_VOL_VELOCITY: if (TM_HAS(LETF_ST_VOL_VELOCITY)) {
              R ev = s_v5 > s_v20 ? s_v5 : s_v20;       // python max(vol_slow, vol_fast)
              if (isnan_(ev) || ev < (R)0.001) ev = (R)0.5;
              alloc = clipr(M::div_(st.p[0], ev), (R)0, (R)1);
            } break;
            case LETF_ST_VOL_MOM: if (TM_HAS(LETF_ST_VOL_MOM)) {
              R v = s_v20; if (isnan_(v) || v < (R)0.001) v = (R)0.5;
              R al = M::div_(st.p[0], v);
              if (!isnan_(s_vm) && s_vm < (R)-0.10) al *= (R)1.2;
              if (!isnan_(s_vm) && s_vm > (R)0.10) al *= (R)0.7;
              alloc = clipr(al, (R)0, (R)1);
            } break;
            case LETF_ST_SKEW_CONVEX: if (TM_HAS(LETF_ST_SKEW_CONVEX)) {
              R v = s_v20; if (isnan_(v) || v < (R)0.001) v = (R)0.5;
              R al = M::div_(st.p[0], v);
              if (!isnan_(s_sk) && s_sk > (R)0) al = M::pow_(al, (R)1.3);
              else if (!isnan_(s_sk) && s_sk < (R)-0.5) al *= (R)0.5;
              alloc = clipr(al, (R)0, (R)1);
            } break;
            case LETF_ST_META_ENSEMBLE: if (TM_HAS(LETF_ST_META_ENSEMBLE)) {
              R dv = s_dv20; if (isnan_(dv) || dv < (R)0.001) dv = (R)0.25;
              R al = M::div_(st.p[0], dv);
              if (P_prev > s_sma200) al = M::pow_(al, (R)1.2);
              if (s_v5 > (R)1.5 * s_v20) al *= (R)0.5;
              alloc = clipr(al, (R)0, (R)1);
            } break;
            case LETF_ST_REGIME_ASYMMETRIC: if (TM_HAS(LETF_ST_REGIME_ASYMMETRIC)) {
              R vr = M::div_(s_v5, s_v60);
              bool crisis = (vix > st.p[2]) || (vr > st.p[3]);
              R tv = crisis ? st.p[1] : st.p[0];
              alloc = (isnan_(s_v5) || s_v5 < (R)0.001) ? (R)0.5 : clipr(M::div_(tv, s_v5), (R)0, (R)1.2);
            } break;
            case LETF_ST_SKEW_KELLY: if (TM_HAS(LETF_ST_SKEW_KELLY)) {
              R vol = s_v20, dv = s_dv20, sk = s_sk, mr = s_mean60;
              if (isnan_(vol) || vol < (R)0.001) vol = (R)0.25;
              if (isnan_(dv) || dv < (R)0.001) dv = vol * (R)0.6;
              if (isnan_(sk)) sk = 0;
              if (isnan_(mr)) mr = (R)0.08;
              R ev, pen;
              if (sk < (R)-0.5) { ev = dv * (R)1.5; pen = (R)0.6; }
              else if (sk < (R)0) { ev = dv * (R)1.2; pen = (R)0.8; }
              else if (sk > (R)0.5) { ev = vol * (R)0.9; pen = (R)1.2; }
              else { ev = vol; pen = (R)1; }
              R ks = (R)1;
              if (ev > (R)0.01) { ks = M::div_(mr - (R)0.03, ev * ev); ks = ks * st.p[3]; ks = clipr(ks, (R)0.2, (R)2.0); }
              alloc = clipr(M::div_(st.p[0], ev) * pen * M::div_(ks, (R)1.5), (R)0, (R)1.5);
            } break;
            case LETF_ST_MOM_VOL_CONV: if (TM_HAS(LETF_ST_MOM_VOL_CONV)) {
              R mom = s_mom, vf = s_v10, vs = s_v60;
              if (isnan_(mom)) mom = 0;
              if (isnan_(vf) || vf < (R)0.001) vf = (R)0.30;
              if (isnan_(vs) || vs < (R)0.001) vs = (R)0.25;
              R mmul = mom > st.p[4] ? (R)1.3 : (mom > (R)0 ? (R)1.0 : (R)0.5);
              R ratio = M::div_(vf, vs), vm, ev;
              if (ratio < (R)0.8) { vm = (R)1.2; ev = vf; } else if (ratio > (R)1.3) { vm = (R)0.6; ev = vf; } else { vm = (R)1.0; ev = vs; }
              alloc = clipr(M::div_(st.p[0] * (mmul * vm), ev), (R)0, (R)1.5);
            } break;
            case LETF_ST_CONVICTION: if (TM_HAS(LETF_ST_CONVICTION)) {
              R mom = s_mom, dv = s_dv20, vf = s_v5, vs = s_v60;
              if (isnan_(mom)) mom = 0;
              if (isnan_(dv) || dv < (R)0.001) dv = (R)0.20;
              if (isnan_(vf) || vf < (R)0.001) vf = (R)0.30;
              if (isnan_(vs) || vs < (R)0.001) vs = (R)0.25;
              R ms = mom > (R)0.15 ? (R)1.0 : (mom > (R)0.05 ? (R)0.7 : (mom > (R)0 ? (R)0.4 : (R)0));
              R ts = P_prev > s_sma100 ? (R)0.5 : (R)0;
              R ratio = M::div_(vf, vs);
              R vsc = ratio < (R)0.9 ? (R)0.5 : (ratio < (R)1.2 ? (R)0.3 : (R)0);
              R conv = ms + ts + vsc;
              alloc = clipr(M::div_(st.p[0], dv) * ((R)0.3 + (conv * (R)0.55)), (R)0, (R)1.5);
              thresh = st.p[4]; sticky = true;
            } break;
            default: break;
          }
          if (EXACT || JOURNAL) {
            if (skip) { sl.dd.step(sl.W, (R)1); return; }
            R prev_alloc = sl.aux;
            R turnover = M::abs_(alloc - prev_alloc), cost = 0;
            if (turnover > thresh) {
              sl.trades++;
              if (JOURNAL && st.taxable) jlog(sj[s], 0, s, i, st.asset, alloc, sl.W, (R)(cp[st.asset] * 100.0));
              cost = turnover * (R)0.0003;
              sl.aux = alloc;
            } else {
              if (sticky) alloc = prev_alloc; else sl.aux = alloc;
            }
            r_strat = cashleg ? (alloc * r_asset) + (((R)1 - alloc) * cash) : r_asset * alloc;
            if (p.apply_costs) r_strat -= cost;
          } else {
            // same decisions without divergent branches (fp32 throughput builds keep no journal)
            const R prev_alloc = sl.aux;
            if (skip) alloc = prev_alloc;                      // strategy.py: the day is skipped, allocation state unchanged
            const R turnover = M::abs_(alloc - prev_alloc);
            const bool traded = turnover > thresh;
            sl.trades += traded ? 1 : 0;
            const R cost = traded ? turnover * (R)0.0003 : (R)0;
            if (sticky) alloc = traded ? alloc : prev_alloc;
            sl.aux = alloc;
            r_strat = cashleg ? (alloc * r_asset) + (((R)1 - alloc) * cash) : r_asset * alloc;
            if (p.apply_costs) r_strat -= cost;
            if (skip) r_strat = (R)0;
          }
        }
        growth = (R)1 + r_strat;
        if (EXACT) sl.W = sl.W * (1.0 + (double)r_strat);
        else sl.W = ::fma(sl.W, (double)r_strat, sl.W);
        sl.dd.step(sl.W, growth);
      };
      if (HAS_ACTIVE && i >= 1) {         // active strategies leave equity[0] = W0
        if (LIST != 0) for_typed_slots<LIST>(slot_body, std::make_integer_sequence<int, NS>{});
        else {
#pragma unroll(NS <= 8 ? NS : 1)
          for (int s = 0; s < NS; s++) slot_body(s, std::integral_constant<int, -1>{}, std::integral_constant<int, -1>{});
        }
      }
      if (TAPEIO && ka.log_curve && !ka.trial_ids && ka.log_trial == k && ka.log_strategy >= 0 && ka.log_strategy < p.n_strat) {
        // equity curve of one strategy (tape-capable builds only: their slot loop is rolled, slot[] lives in local memory)
        const int s = ka.log_strategy, ty = p.strat[s].type;
        double w = slot[s].W;
        if (ty == LETF_ST_BENCHMARK_TLT) w = (double)p.W0 * cp_tlt;
        else if (ty == LETF_ST_BENCHMARK) {
          w = (double)p.W0;
#pragma unroll
          for (int a = 0; a < NA; a++) if (p.act_id[a] == p.strat[s].asset) w = (double)p.W0 * cp[a];
        }
        ka.log_curve[i] = w;
      }
    };
    if (EXACT) {
      for (int i = 0; i < T; i++) day_step(i, std::false_type{});
    } else {
      const int warm = T < kSteady ? T : kSteady;
      for (int i = 0; i < warm; i++) day_step(i, std::false_type{});
      for (int i = warm; i < T; i++) day_step(i, std::true_type{});
    }

    // ---- per-trial outputs (engine.py:578-603)
    if (MARKET == LETF_MARKET_NORMAL) {       // constant VIX, regime 0
      if (T > 0) { vmin = vmax = vix; if (!isfinite_(vix) || vix <= (R)0) flags |= LETF_FLAG_BAD_VIX; }
    }
    if (!EXACT) {
      // fp32 mode defers the per-day validity checks: non-finite values and exhausted products are sticky
      if (chk != (R)0) flags |= LETF_FLAG_NONFINITE_RET;
      for (int a = 0; a < NA; a++) {
        bool on = FSER ? false : (MARKET == kMarketSeries ? ka.series.asset_ret[a] != nullptr : p.act_on[a] != 0);
        if (on && T > 0 && !(cp[a] > 0.0 && isfinite_(cp[a]))) bad = true;
      }
      if (p.bench_tlt && T > 0 && !(cp_tlt > 0.0 && isfinite_(cp_tlt))) bad = true;
    }
    if (T == 0 || bad || !(vmin >= (R)5) || !(vmax <= (R)120)) flags |= LETF_FLAG_REPAIRED;
    const int64_t N = ka.out.n_trials;
#pragma unroll(LIST != 0 || NS <= 8 ? NS : 1)
    for (int s = 0; s < NS; s++) {
      // (`continue`, not `break`: with a run-time trip count the unrolled loop gets a rolled remainder loop, whose
      //  dynamic slot[s] forces the whole slot array - wealth, drawdown, trade counts - through local memory every day)
      if (s >= p.n_strat) continue;
      const int ty = p.strat[s].type;
      if (FSER && (ty == LETF_ST_BENCHMARK || ty == LETF_ST_BENCHMARK_TLT)) continue;       // written by the producer
      if (ROUTED && p.strat[s].route < 0) continue;
      const int64_t row = ROUTED && p.strat[s].route > 0 ? p.strat[s].route - 1 : s;
      double fw = T > 0 ? slot[s].W : 0.0;
      double md = T > 1 ? slot[s].dd.value() : 0.0;
      if (ty == LETF_ST_BENCHMARK || ty == LETF_ST_BENCHMARK_TLT) {
        double c = 1.0; Drawdown<R> d = ddT;
        if (ty == LETF_ST_BENCHMARK_TLT) c = cp_tlt;
        else {
#pragma unroll
          for (int a = 0; a < NA; a++) if (p.act_on[a] && p.act_id[a] == p.strat[s].asset) { c = cp[a]; d = ddA[a]; }
        }
        fw = T > 0 ? (double)p.W0 * c : 0.0;
        md = T > 1 ? d.value() : 0.0;
      }
      ka.out.final_wealth[row * N + col] = fw;
      ka.out.max_dd[row * N + col] = md;
      ka.out.num_trades[row * N + col] = slot[s].trades;
      if (ka.out.journal_count) ka.out.journal_count[row * N + col] = JOURNAL ? sj[s].cnt : 0;
      if (ka.out.journal_volume) ka.out.journal_volume[row * N + col] = JOURNAL ? sj[s].vol : 0.0;
    }
#if defined(__CUDA_ARCH__)
    if (FSER) cp_async_wait<0>();              // nothing of this trial's staging may still be in flight when the next one starts
#endif
    if (FSER) return;                          // path metadata and validity flags are the producer's
    if (ka.out.start_offset) ka.out.start_offset[col] = so;
    if (ka.out.flags) ka.out.flags[col] = flags;
    if (ka.out.regime_days) { ka.out.regime_days[col] = T - reg_days1; ka.out.regime_days[N + col] = reg_days1; }   // NORMAL: reg_days1 stays 0
    if (ka.out.jump_days) ka.out.jump_days[col] = jump_days;
  }

  // VIX recursion (engine.py:100-115)
  LETF_HD R vix_step(int g, int reg, R vix_prev, R spy_ret, R noise) const {
    if (g == 0) return p.vix_base[reg];
    const DevVix<R>& v = p.vix[reg];
    R shock = M::div_(rmax(-spy_ret, (R)0), v.denom);
    R vj = v.jump_scale * rmax((R)0, shock - v.jump_thr);
    R cand = v.phi * vix_prev + v.drift + vj + ((R)0 + v.noise_std * noise);
    return cand > (R)10 ? cand : (R)10;
  }
};

#undef TM_HAS

}  // namespace letf
