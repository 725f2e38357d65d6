// letf_api.cu — the C-ABI of include/letf_b200.h: engine lifecycle, kernel selection, launches,
// host-buffer convenience entry points.  No CPU compute path exists in this library: every
// entry point that produces path results launches a device kernel or fails.
//
// Reference seam replaced:
//   letf/mc_runner.py:155-165          loky fan-out                  -> letf_run / letf_run_host
//   letf/simulation/engine.py:423-619  simulate_single_path_fixed    -> the fused path kernel
//   letf/strategy.py:55                run_strategy_fixed(df, ...)   -> letf_run_strategies
//   letf/trade.py:83-91                TradeJournal.get_full_trades  -> letf_trade_list_host
//   letf/reporting.py:358-376          percentile / win-rate inputs  -> letf_reduce
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <vector>
#include <chrono>
#include <thread>
#include "letf_kernels.cuh"
#include "letf_series_prep.cuh"

namespace letf {

// ---- kernel registry (filled by the static registrars in inst_*.cu) -------------------------
static std::vector<KernelEntry>& registry() {
  static std::vector<KernelEntry> r;
  return r;
}
void register_kernel(const KernelEntry& e) { registry().push_back(e); }
int kernel_count() { return (int)registry().size(); }
const KernelEntry* kernel_at(int i) { return &registry()[i]; }

cudaError_t launch_reduce(const LetfOutputs& out, const LetfSummary& sm, int64_t n, int n_strat, double w0,
                          double years, int sm_count, cudaStream_t st);
cudaError_t launch_summary_reset(const LetfSummary& sm, int n_strat, cudaStream_t st);
cudaError_t launch_summary_combine(const void* gathered, int world, int n_strat, void* dst, cudaStream_t st);
cudaError_t launch_series_records(const SeriesPrepArgs& a, cudaStream_t st);

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return LETF_ERR_CUDA;
}
#define CK(call)                                         \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) return fail_cuda(_e, #call);  \
  } while (0)

constexpr uint32_t kNeedRingP = (1u << LETF_ST_SMA) | (1u << LETF_ST_SMA_BAND) | (1u << LETF_ST_COMPOSITE) |
                                (1u << LETF_ST_ADAPTIVE_VOL) | (1u << LETF_ST_CONVEX_VOL) |
                                (1u << LETF_ST_META_ENSEMBLE) | (1u << LETF_ST_CONVICTION);
constexpr uint32_t kNeedRingS = (1u << LETF_ST_MOM_VOL_CONV) | (1u << LETF_ST_CONVICTION);

static inline size_t align256_(size_t x) { return (x + 255) & ~(size_t)255; }

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// page-locked host staging: device -> pinned at full PCIe rate, then a plain memcpy into the caller's
// (pageable) arrays; allocating / freeing pinned memory per call costs milliseconds and synchronises
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace letf

using namespace letf;

struct LetfEngine {
  LetfParams params;
  int device = 0, precision = 0, sm_count = 0;
  int32_t* dev_dur[2] = {nullptr, nullptr};
  std::vector<int32_t> host_dur[2];
  uint32_t type_mask = 0;
  bool any_taxable = false;
  std::string pin_kernel;                    // tuning aids, read from the environment when the engine is created:
  bool no_jpair = false;                     // LETF_KERNEL=<instantiation>, LETF_NO_JPAIR=1, LETF_NO_SPLIT=1
  bool no_split = false;
  bool no_generic_split = false;             // LETF_NO_GENERIC_SPLIT=1: untyped strategy lists keep the fused builds (A/B runs)
  bool trace_host = false;
  int force_split_waves = 0;                 // LETF_SPLIT_WAVES=<1..4>: producer waves per chunk of the split launch (tests: chunk seams)
  int64_t force_series_chunk = 0;            // LETF_SERIES_CHUNK=<paths>: paths per chunk of the materialised-path mode
  bool no_routed_split = false;              // LETF_NO_ROUTED_SPLIT=1: subsets of the default sweep skip the routed 19-rule kernel
  DevBuf rec;                                // per-day records of the split strategy pipeline
  DevBuf rec_cash;                           // + the frame's Cash_Ret column in the materialised-path mode
  int64_t split_key = 0; int split_waves = 0; // its chunk size in producer waves, remembered per shape (launch_split)
  int64_t series_key = 0, series_cap = 0;    // same for the materialised-path mode (launch_series_split), in paths
  DevBuf spell;                              // regime spells of the pair kernels
  // legacy_hybrid bootstrap tables on the device, in the engine's precision (letf_params.cuh build_boot_tables)
  void* boot_blob = nullptr;
  DevBoot<double> boot64;
  DevBoot<float> boot32;
  std::vector<double> host_hist[3];
  std::vector<int32_t> host_pstart[2][2];
  std::vector<double> host_pret[2][2];
  DevBuf xscr, ringP, ringS, ringA;          // per-thread scratch of the fused kernel
  DevBuf stage_tape, stage_out, stage_log;   // device staging of the *_host entry points
  PinBuf pin_out;                            // pinned mirror of stage_out
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  LetfLaunchInfo info;
  std::mutex mu;
};

namespace letf {

struct LogArgs {
  const int64_t* trial_ids = nullptr;       // device: batch replay (letf_trade_lists_host)
  const int32_t* strategies = nullptr;      // device: its per-item strategies (letf_trade_lists_multi_host)
  int64_t trial = -1; int strategy = -1; int cap = 0;
  int32_t* count = nullptr; int32_t* day = nullptr; int32_t* asset = nullptr; int32_t* action = nullptr;
  double* shares = nullptr; double* price = nullptr;
  double* curve = nullptr;                  // device [T]: equity curve of (trial, strategy)
  double* market = nullptr;                 // device [T][3]: generated SPY / QQQ / TLT returns of `trial`
};

// typed slots: the strategy list must be exactly the one compiled in
static bool list_matches(const KernelEntry* k, const LetfEngine* e) {
  const int n_strat = e->params.n_strategies;
  bool same = slot_list_count(k->list) == n_strat;
  for (int s = 0; same && s < n_strat; s++) {
    same = slot_list_type(k->list, s) == e->params.strategies[s].type;
    const int as = slot_list_asset(k->list, s);
    if (same && as >= 0) same = e->params.strategies[s].asset == as;
    if (same && e->params.strategies[s].type == LETF_ST_COMPOSITE) same = e->params.strategies[s].asset2 == 4;
  }
  return same;
}

// Split strategy pipeline: the market + ETF layers of one antithetic pair per thread (letf_jfast.cuh, EMIT build) write
// per-day records, a rule kernel (Trial::run in record mode, one trial per thread) consumes them.  Admitted: fp32,
// Philox draws, no journal / tape plumbing, the institutional shape the pair kernel serves (joint-t with or without moment
// matching, antithetic pairs, stress + funding models).  Three tiers, tried in this order:
//  * typed (find_split): the list is exactly a compiled one, SPY simulated and QQQ not (compacted producer slots, the SPY
//    price is the last record field);
//  * routed (find_routed_split): the list is a subset of a compiled one - its rule kernel in the routed build, the other
//    slots switched off, output rows routed; the producer simulates the compiled list's series;
//  * generic (find_generic_split): any other list - the uncompacted producer writes a field for each of the six series
//    (zeros outside the asset mask), a dynamic-slot rule kernel reads them.
// No rule kernel is launched when every strategy is buy&hold (the producer writes those rows).
struct SplitPlan {
  const KernelEntry* prod = nullptr; const KernelEntry* cons = nullptr; int nact = 0; bool rules = true;
  // routed: the requested strategies sit in slots of a larger typed list (slot_of[k] = slot of requested strategy k)
  bool routed = false; int slot_of[LETF_MAX_STRATEGIES]; int n_slots = 0; int mask = 0;
};
static bool split_shape_ok(const LetfEngine* e, int rng_mode, bool need_io, int64_t first) {
  const LetfParams& s = e->params;
  if (e->precision != LETF_PREC_F32 || need_io || rng_mode != LETF_RNG_PHILOX || e->no_split) return false;
  return s.market_mode == LETF_MARKET_JOINT_T && s.antithetic_pairs && s.institutional &&
         s.has_stress && s.has_funding && s.sim_days >= 2;
}
static bool find_split(const LetfEngine* e, int rng_mode, bool need_io, int64_t first, SplitPlan* plan) {
  const LetfParams& s = e->params;
  if (!split_shape_ok(e, rng_mode, need_io, first)) return false;
  const int mask = s.asset_mask & 0x3F;
  if (!((mask >> 4) & 1) || ((mask >> 5) & 1)) return false;
  int nact = 0, k10 = 0;
  bool tedf = false;
  for (int a = 0; a < kAssets; a++) {
    if (!((mask >> a) & 1)) continue;
    const int fk = df_fast_k(s.assets[a].te_df);
    if ((fk != 5 && fk != 10) || s.assets[a].te_clip > 0.0) tedf = true;      // calibrated tracking_residual_model: run-time sampler build
    if (fk == 10) k10 |= 1 << nact;
    nact++;
  }
  if (tedf) k10 = 0x80;
  for (int i = 0; i < kernel_count(); i++) {
    const KernelEntry* k = kernel_at(i);
    if (k->precision != LETF_PREC_F32) continue;
    // (first registered candidate wins unless LETF_KERNEL names another one)
    if (k->fj == (kSplitProducer | fast_joint_code(nact, k10)) && k->ns >= s.n_strategies && k->na >= nact && k->na < kAssets &&
        (!plan->prod || e->pin_kernel == k->name)) plan->prod = k;
    if (k->market == kMarketFSeries && k->list && !slot_list_routed(k->list) && (k->am & 0x3F) == mask && k->ns >= s.n_strategies && list_matches(k, e) &&
        (!plan->cons || e->pin_kernel == k->name)) plan->cons = k;
  }
  plan->nact = nact;
  return plan->prod && plan->cons;
}
// any subset of a typed list (slot lists 3 and 4: the default 19-strategy sweep and the config-4 six): every rule of the request goes to the slot
// that compiles its type, buy&hold strategies to the list's buy&hold slots or to the free slots behind it (they cost the
// rule kernel nothing: the producer writes them), every other slot is switched off; the producer simulates the list's series
static bool find_routed_split(const LetfEngine* e, int rng_mode, bool need_io, int64_t first, SplitPlan* plan) {
  const LetfParams& s = e->params;
  if (!split_shape_ok(e, rng_mode, need_io, first) || e->no_generic_split || e->no_routed_split) return false;
  *plan = SplitPlan();
  // (several routed lists can hold a request: the one with the fewest slots is the cheapest rule kernel)
  for (int i = 0; i < kernel_count(); i++) {
    const KernelEntry* k = kernel_at(i);
    if (k->precision != LETF_PREC_F32 || k->market != kMarketFSeries || !slot_list_routed(k->list)) continue;
    if (plan->cons && slot_list_count(k->list) >= slot_list_count(plan->cons->list)) continue;
    const int mask = k->am & 0x3F, nl = slot_list_count(k->list);
    if ((s.asset_mask & 0x3F) & ~mask) continue;
    bool rules = false;
    int slot_of[LETF_MAX_STRATEGIES];
    bool ok = route_request(s, k->list, k->ns, slot_of, &rules);
    for (int a = 0; ok && a < kAssets; a++) if (((mask >> a) & 1) && !(s.assets[a].leverage != 0.0)) ok = false;   // (series outside the request need a layer description)
    if (!ok) continue;
    plan->cons = k; plan->mask = mask; plan->rules = rules; plan->routed = true;
    for (int q = 0; q < s.n_strategies; q++) plan->slot_of[q] = slot_of[q];
    plan->n_slots = nl;
    for (int q = 0; q < s.n_strategies; q++) if (plan->slot_of[q] + 1 > plan->n_slots) plan->n_slots = plan->slot_of[q] + 1;
  }
  if (!plan->cons) return false;
  int nact = 0, k10 = 0;
  bool tedf = false;
  for (int a = 0; a < kAssets; a++) {
    if (!((plan->mask >> a) & 1)) continue;
    const int fk = df_fast_k(s.assets[a].te_df);
    if ((fk != 5 && fk != 10) || s.assets[a].te_clip > 0.0) tedf = true;
    if (fk == 10) k10 |= 1 << nact;
    nact++;
  }
  if (tedf) k10 = 0x80;
  for (int i = 0; i < kernel_count(); i++) {
    const KernelEntry* k = kernel_at(i);
    if (k->precision == LETF_PREC_F32 && k->fj == (kSplitProducer | fast_joint_code(nact, k10)) && k->ns >= plan->n_slots &&
        k->na >= nact && k->na < kAssets && (!plan->prod || e->pin_kernel == k->name)) plan->prod = k;
  }
  plan->nact = nact;
  return plan->prod != nullptr;
}
static bool find_generic_split(const LetfEngine* e, int rng_mode, bool need_io, int64_t first, SplitPlan* plan) {
  const LetfParams& s = e->params;
  if (!split_shape_ok(e, rng_mode, need_io, first) || e->no_generic_split) return false;
  const int mask = s.asset_mask & 0x3F;
  // default tracking-error dof (config.py:62-135: t(5) for the leveraged series, t(10) for SPY / QQQ) or the run-time sampler build
  bool tedf = false;
  for (int a = 0; a < kAssets; a++) {
    if (!((mask >> a) & 1)) continue;
    if (df_fast_k(s.assets[a].te_df) != (a >= 4 ? 10 : 5) || s.assets[a].te_clip > 0.0) tedf = true;
  }
  const int code = kSplitProducer | fast_joint_code(kAssets, tedf ? 0x80 : 0x30);
  *plan = SplitPlan();
  for (int i = 0; i < kernel_count(); i++) {
    const KernelEntry* k = kernel_at(i);
    if (k->precision != LETF_PREC_F32) continue;
    if (k->fj == code && k->na == kAssets && k->ns >= s.n_strategies && !plan->prod) plan->prod = k;
    // (the build with the fewest slots that holds the list: eight slots unroll into registers, thirty-two live in local memory)
    if (k->market == kMarketFSeries && !k->list && (k->am & 0x3F) == 0x3F && k->ns >= s.n_strategies &&
        !(e->type_mask & ~k->tm) && (!plan->cons || k->ns < plan->cons->ns || e->pin_kernel == k->name) &&
        !(plan->cons && e->pin_kernel == plan->cons->name)) plan->cons = k;
  }
  plan->nact = kAssets;
  plan->rules = false;
  for (int k = 0; k < s.n_strategies; k++)
    if (s.strategies[k].type != LETF_ST_BENCHMARK && s.strategies[k].type != LETF_ST_BENCHMARK_TLT) plan->rules = true;
  return plan->prod && plan->cons;
}

static const KernelEntry* select_kernel(const LetfEngine* e, int market, bool need_io, int n_strat, uint32_t tmask,
                                       int n_assets) {
  const KernelEntry* best = nullptr;
  const char* pin = e->pin_kernel.c_str();      // LETF_KERNEL, read once per engine (letf_create)
  for (int i = 0; i < kernel_count(); i++) {
    const KernelEntry* k = kernel_at(i);
    if (k->precision != e->precision || k->market != market) continue;
    if (need_io && !k->tapeio) continue;
    if (k->ns < n_strat || (tmask & ~k->tm) || k->na < n_assets) continue;
    // antithetic-pair kernels serve exactly one shape (letf_jfast.cuh); LETF_NO_JPAIR=1 keeps the generic lean build (A/B runs)
    if (k->fj && (k->fj != fast_joint_shape(e->params, k->na) || e->no_jpair)) continue;      // (producers carry kSplitProducer: never equal)
    if (k->am) {        // compile-time asset mask: same series, default tracking-error dof (config.py:62-135)
      bool ok = (k->am & 0x3F) == e->params.asset_mask;
      if (!(k->am & 0x40))          // bit 6: tracking-error dof read at run time (calibrated tracking_residual_model)
        for (int a = 0; ok && a < kAssets; a++)
          if ((k->am >> a) & 1) ok = df_fast_k(e->params.assets[a].te_df) == (a >= 4 ? 10 : 5);
      if (!ok) continue;
    }
    if (k->list && !list_matches(k, e)) continue;
    // tuning aid: LETF_KERNEL=<instantiation name> pins an eligible build
    if (pin && *pin && !strcmp(pin, k->name)) return k;
    if (!best) { best = k; continue; }
    // prefer the leanest build: fewer strategy slots, fewer rule types, no tape plumbing
    // typed-slot builds first, then fewer slots / series / rules, no tape plumbing
    int a = (k->list ? 0 : 100000) + k->ns * 64 + k->na * 16 + __builtin_popcount(k->tm) * 2 + k->tapeio + ((k->am & 0x40) ? 1 : 0) -
            (k->fj ? 1 : 0);
    int b = (best->list ? 0 : 100000) + best->ns * 64 + best->na * 16 + __builtin_popcount(best->tm) * 2 + best->tapeio +
            ((best->am & 0x40) ? 1 : 0) - (best->fj ? 1 : 0);
    if (a < b) best = k;
  }
  return best;
}

template <typename R, int NS, int NA>
static int launch_typed(LetfEngine* e, const KernelEntry* ke, cudaStream_t st, uint64_t seed, int64_t first,
                        int64_t n, int rng_mode, const LetfTape* tape, const LetfOutputs* out,
                        const LetfSeries* series, const LogArgs& lg) {
  cudaFuncAttributes fa;
  int max_blocks = 0;
  CK(ke->attrs(&fa, &max_blocks));
  if (max_blocks < 1) return fail(LETF_ERR_CUDA, std::string("kernel ") + ke->name + " cannot be resident on this device");
  static thread_local KArgs<R, NS, NA> ka;
  memset(&ka, 0, sizeof(ka));
  convert_params<R, NS, NA>(e->params, e->dev_dur, ka.p,
                            sizeof(R) == 8 ? (const DevBoot<R>*)(const void*)&e->boot64 : (const DevBoot<R>*)(const void*)&e->boot32);
  // persistent grid: every resident thread owns ceil(n / nthreads) work items; shrink the grid so that
  // all waves are full (equal-length trials: a partial last wave costs a whole trial time).  A work item is
  // one trial, or `upt` interleaved trials in the straight-line fast path (letf_fast.cuh).
  const int kBlock = ke->block;
  const int upt = (ke->upt > 1 && fast_pattern_compiled<NA>(ka.p.fast_layers) && rng_mode == LETF_RNG_PHILOX && !series) ? ke->upt : 1;
  // pair kernels: one work item per antithetic pair the trial range touches
  const int64_t items = ke->fj ? ((first + n - 1) >> 1) - (first >> 1) + 1 : (n + upt - 1) / upt;
  const int64_t resident = (int64_t)e->sm_count * max_blocks * kBlock;
  const int64_t waves = (items + resident - 1) / resident;
  const int64_t per_wave = (items + waves - 1) / waves;
  int64_t grid = (per_wave + kBlock - 1) / kBlock;
  if (grid < 1) grid = 1;
  const int64_t nthreads = grid * kBlock;
  if ((int64_t)kRingP * nthreads >= ((int64_t)1 << 32)) return fail(LETF_ERR_INVALID, "grid too large for 32-bit ring offsets");

  if (series) { ka.series = *series; ka.p.T = series->sim_days; ka.p.buffer = 0; ka.p.random_start = 0; }
  if (tape) ka.tape = *tape;
  ka.out = *out;
  ka.seed = seed; ka.first_trial = first; ka.n_trials = n; ka.tape_first = 0;
  ka.rng_mode = rng_mode; ka.nthreads = nthreads;
  ka.trial_ids = lg.trial_ids; ka.log_strategies = lg.strategies;
  ka.log_trial = lg.trial; ka.log_strategy = lg.strategy; ka.log_cap = lg.cap;
  ka.log_count = lg.count; ka.log_day = lg.day; ka.log_asset = lg.asset; ka.log_action = lg.action;
  ka.log_shares = lg.shares; ka.log_price = lg.price; ka.log_curve = lg.curve; ka.log_market = lg.market;

  const int G = e->params.sim_days + e->params.buffer_days;
  size_t scratch = 0;
  // (the pair kernels are two-pass with or without moment matching: pass 2 reads pass 1's day records either way)
  if (((ke->market == LETF_MARKET_JOINT_T && (e->params.moment_matching || ke->fj)) || ke->market == LETF_MARKET_BOOTSTRAP) && G > 0) {
    size_t b = ((size_t)G * 4 + (size_t)((G + 3) & ~3)) * (size_t)nthreads * sizeof(R);   // day vectors + the IRX-normal column (letf_core.cuh AUXSCR)
    if (ke->fj) {
      b = (size_t)G * sizeof(JF8) * (size_t)nthreads;                                      // one 32-byte record per day and pair (letf_jfast.cuh)
      const size_t sb = (size_t)(G + 1) * 4 * (size_t)nthreads;                            // + the pair's regime spells
      cudaError_t ce2 = e->spell.reserve(sb);
      if (ce2 != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "regime spell scratch: " + std::string(cudaGetErrorString(ce2))); }
      ka.spell_scr = (uint32_t*)e->spell.p; scratch += sb;
    }
    cudaError_t ce = e->xscr.reserve(b);
    if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "moment-matching scratch: " + std::string(cudaGetErrorString(ce))); }
    ka.xscr = (R*)e->xscr.p; scratch += b;
  }
  if (ke->tm & kNeedRingP) { size_t b = (size_t)kRingP * nthreads * sizeof(R); CK(e->ringP.reserve(b)); ka.ringP = (R*)e->ringP.p; scratch += b; }
  if (ke->tm & kNeedRingS) { size_t b = (size_t)kRingS * nthreads * sizeof(R); CK(e->ringS.reserve(b)); ka.ringS = (R*)e->ringS.p; scratch += b; }
  if (ke->tm & kVolTypes) { size_t b = (size_t)kRingA * nthreads * sizeof(R); CK(e->ringA.reserve(b)); ka.ringA = (R*)e->ringA.p; scratch += b; }

  CK(cudaEventRecord(e->ev0, st));
  CK(ke->launch(&ka, (int)grid, st));
  CK(cudaEventRecord(e->ev1, st));
  e->timed = true;
  e->info.grid = (int)grid; e->info.block = kBlock; e->info.regs_per_thread = fa.numRegs;
  e->info.kernel_launches = 1; e->info.scratch_bytes = (int64_t)scratch; e->info.sm_count = e->sm_count;
  snprintf(e->info.kernel, sizeof(e->info.kernel), "%s", ke->name);
  return LETF_OK;
}

// full-wave persistent grid for `items` equal-length work items
static int64_t full_wave_grid(int64_t items, int sm_count, int max_blocks, int block) {
  const int64_t resident = (int64_t)sm_count * max_blocks * block;
  const int64_t waves = (items + resident - 1) / resident;
  const int64_t per_wave = (items + waves - 1) / waves;
  int64_t grid = (per_wave + block - 1) / block;
  return grid < 1 ? 1 : grid;
}
static LetfOutputs out_columns(const LetfOutputs& o, int64_t off) {
  LetfOutputs r = o;             // [S][N] arrays: column j of every row is (pointer + j), the pitch N = n_trials stays
  r.final_wealth += off; r.max_dd += off; r.num_trades += off;
  if (r.journal_count) r.journal_count += off;
  if (r.journal_volume) r.journal_volume += off;
  if (r.start_offset) r.start_offset += off;
  if (r.flags) r.flags += off;
  if (r.regime_days) r.regime_days += off;
  if (r.jump_days) r.jump_days += off;
  return r;
}

template <int NSC, int NAP>
static int launch_split(LetfEngine* e, const SplitPlan& plan, cudaStream_t st, uint64_t seed, int64_t first, int64_t n,
                        const LetfOutputs* out) {
  constexpr int NSP = 32;                              // producer instantiations (inst_f32_split_producers.cu: NAP = 4 compacted
                                                       // slots; inst_f32_split_generic_prod*.cu: NAP = 6, slot = asset)
  const KernelEntry *kp = plan.prod, *kc = plan.cons;
  const bool trace = e->trace_host;           // LETF_TRACE_HOST=1: host time of each step of this function on stderr (tuning aid)
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[split %s %.3f ms] ", what, std::chrono::duration<double, std::milli>(t - tp0).count());
    tp0 = t;
  };
  cudaFuncAttributes fap, fac;
  int mbp = 0, mbc = 0;
  CK(kp->attrs(&fap, &mbp));
  CK(kc->attrs(&fac, &mbc));
  if (mbp < 1 || mbc < 1) return fail(LETF_ERR_CUDA, "split pipeline kernels cannot be resident on this device");
  lap("attrs");
  static thread_local KArgs<float, NSP, NAP> ka_p;
  static thread_local KArgs<float, NSC, kAssets> ka_c;
  memset(&ka_p, 0, sizeof(ka_p)); memset(&ka_c, 0, sizeof(ka_c));
  if (plan.routed) {
    static thread_local LetfParams rp;
    rp = routed_params(e->params, plan.cons->list, plan.slot_of, plan.n_slots, plan.mask);
    convert_params<float, NSP, NAP>(rp, e->dev_dur, ka_p.p, (const DevBoot<float>*)(const void*)&e->boot32);
    convert_params<float, NSC, kAssets>(rp, e->dev_dur, ka_c.p, (const DevBoot<float>*)(const void*)&e->boot32);
    route_slots(ka_p.p, e->params, plan.slot_of, plan.n_slots); route_slots(ka_c.p, e->params, plan.slot_of, plan.n_slots);
  } else {
    convert_params<float, NSP, NAP>(e->params, e->dev_dur, ka_p.p, (const DevBoot<float>*)(const void*)&e->boot32);
    convert_params<float, NSC, kAssets>(e->params, e->dev_dur, ka_c.p, (const DevBoot<float>*)(const void*)&e->boot32);
  }
  lap("params");
  const int T = e->params.sim_days, G = T + e->params.buffer_days, NF = plan.rules ? plan.nact + 2 : 0;   // (no records without a rule kernel)
  // chunk of trials: whole producer waves (two trials per thread), up to four, as many as fit in 60 % of the free device
  // memory beside the producer's per-wave scratch.  The free-memory query is made once per (horizon, fields, wave) shape and
  // remembered: cudaMemGetInfo costs 0.2 - 70 ms of host time with tens of GB allocated (measured), on every call's critical path
  const int64_t wave_trials = 2 * (int64_t)e->sm_count * mbp * kp->block;
  const double per_trial = (double)T * NF * 4;
  const int64_t shape_key = ((int64_t)T << 32) ^ ((int64_t)NF << 24) ^ wave_trials;
  int64_t waves_fit = e->split_waves;
  if (e->split_key != shape_key || waves_fit < 1) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const double wave_scratch = ((double)G * sizeof(JF8) + (double)(G + 1) * 4) * (double)(wave_trials / 2);
    const double held = (double)e->xscr.cap + (double)e->spell.cap;
    const double budget = 0.6 * ((double)free_b + (double)e->rec.cap - (wave_scratch > held ? wave_scratch - held : 0.0));
    waves_fit = per_trial > 0 ? (int64_t)(budget / (per_trial * (double)wave_trials)) : 4;
    if (waves_fit < 1) return fail(LETF_ERR_NOMEM, "split pipeline: one wave of per-day records does not fit in device memory");
    if (waves_fit > 4) waves_fit = 4;
    e->split_key = shape_key; e->split_waves = (int)waves_fit;
  }
  if (e->force_split_waves >= 1 && e->force_split_waves < waves_fit) waves_fit = e->force_split_waves;
  int64_t cap = 0;
  for (;;) {
    cap = wave_trials * waves_fit;
    if (cap > ((n + 1) & ~(int64_t)1)) cap = (n + 1) & ~(int64_t)1;
    cudaError_t ce = e->rec.reserve((size_t)(cap + 2) * T * NF * 4);
    if (ce == cudaSuccess) break;
    cudaGetLastError();
    // (the remembered chunk size no longer fits, e.g. the caller allocated in between: smaller chunks, asked again next call)
    e->split_key = 0;
    if (waves_fit <= 1) return fail(LETF_ERR_NOMEM, "split pipeline records: " + std::string(cudaGetErrorString(ce)));
    waves_fit /= 2;
  }
  lap("reserve");
  ka_p.seed = ka_c.seed = seed; ka_p.rng_mode = ka_c.rng_mode = LETF_RNG_PHILOX;
  // record columns are pair-aligned: column 0 is the even twin of the chunk's first pair.  A launch that starts on an odd trial
  // (its first pair half outside the range) reads from column 1 on - the rule kernel simply gets the base shifted by one
  // element; two spare columns hold the half pairs at both ends
  const int odd = (int)(first & 1);
  ka_p.rec = (float*)e->rec.p; ka_c.rec = (float*)e->rec.p + odd; ka_p.rec_ncol = ka_c.rec_ncol = cap + 2;
  ka_p.log_trial = ka_c.log_trial = -1;
  ka_c.p.T = T; ka_c.p.buffer = 0; ka_c.p.random_start = 0;
  ka_c.series.sim_days = T;
  size_t scratch = (size_t)(cap + 2) * T * NF * 4;
  int launches = 0;
  int64_t grid_c_last = 0, grid_p_last = 0;
  CK(cudaEventRecord(e->ev0, st));
  for (int64_t c0 = 0; c0 < n; c0 += cap) {
    const int64_t nc = n - c0 < cap ? n - c0 : cap;
    const LetfOutputs oc = out_columns(*out, c0);
    // producer: one work item per antithetic pair
    const int64_t pairs = ((first + c0 + nc - 1) >> 1) - ((first + c0) >> 1) + 1;
    const int64_t grid_p = full_wave_grid(pairs, e->sm_count, mbp, kp->block);
    ka_p.out = oc; ka_p.first_trial = first + c0; ka_p.n_trials = nc; ka_p.nthreads = grid_p * kp->block;
    {
      const size_t b = (size_t)G * sizeof(JF8) * (size_t)ka_p.nthreads;
      cudaError_t ce = e->xscr.reserve(b);
      if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "moment-matching scratch: " + std::string(cudaGetErrorString(ce))); }
      ka_p.xscr = (float*)e->xscr.p;
      const size_t sb = (size_t)(G + 1) * 4 * (size_t)ka_p.nthreads;
      ce = e->spell.reserve(sb);
      if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "regime spell scratch: " + std::string(cudaGetErrorString(ce))); }
      ka_p.spell_scr = (uint32_t*)e->spell.p;
      if (c0 == 0) scratch += b + sb;
    }
    lap("scratch");
    CK(kp->launch(&ka_p, (int)grid_p, st));
    lap("producer");
    launches++; grid_p_last = grid_p;
    if (!plan.rules) continue;                 // buy&hold-only list: the producer wrote every output
    // rule kernel: one trial per thread, columns [0, nc) of the record buffer
    const int64_t grid_c = full_wave_grid(nc, e->sm_count, mbc, kc->block);
    const int64_t nth = grid_c * kc->block;
    if ((int64_t)kRingP * nth >= ((int64_t)1 << 32)) return fail(LETF_ERR_INVALID, "grid too large for 32-bit ring offsets");
    ka_c.out = oc; ka_c.first_trial = first + c0; ka_c.n_trials = nc; ka_c.nthreads = nth; ka_c.series.n_paths = nc;
    if (kc->tm & kNeedRingP) { size_t b = (size_t)kRingP * nth * 4; CK(e->ringP.reserve(b)); ka_c.ringP = (float*)e->ringP.p; if (c0 == 0) scratch += b; }
    if (kc->tm & kNeedRingS) { size_t b = (size_t)kRingS * nth * 4; CK(e->ringS.reserve(b)); ka_c.ringS = (float*)e->ringS.p; if (c0 == 0) scratch += b; }
    if (kc->tm & kVolTypes) { size_t b = (size_t)kRingA * nth * 4; CK(e->ringA.reserve(b)); ka_c.ringA = (float*)e->ringA.p; if (c0 == 0) scratch += b; }
    CK(kc->launch(&ka_c, (int)grid_c, st));
    lap("rules");
    launches++; grid_c_last = grid_c;
  }
  if (trace) fprintf(stderr, "\n");
  CK(cudaEventRecord(e->ev1, st));
  e->timed = true;
  e->info.grid = (int)grid_p_last; e->info.block = kp->block; e->info.regs_per_thread = fap.numRegs;
  e->info.kernel_launches = launches; e->info.scratch_bytes = (int64_t)scratch; e->info.sm_count = e->sm_count;
  if (plan.rules) snprintf(e->info.kernel, sizeof(e->info.kernel), "%s+%s", kp->name + 4, kc->name + 4);
  else snprintf(e->info.kernel, sizeof(e->info.kernel), "%s", kp->name + 4);
  (void)grid_c_last;
  return LETF_OK;
}

// Materialised-path mode through the split pipeline (letf_series_prep.cuh): a bandwidth-bound record pass over the caller's
// series (it also writes the buy&hold strategies, flags and regime day counts), then the rule kernel the strategy list would
// get on simulated paths - typed, routed or generic.  Not taken: fp64, journal / curve / trade-list plumbing, a frame with its
// own Cash_Ret column (the rule kernels read the regime's cash rate), LETF_NO_SPLIT=1.
static bool find_series_split(const LetfEngine* e, const LetfSeries* series, bool need_io, SplitPlan* plan) {
  const LetfParams& s = e->params;
  *plan = SplitPlan();
  if (e->precision != LETF_PREC_F32 || need_io || e->no_split || series->sim_days < 1) return false;
  const int mask = s.asset_mask & 0x3F;
  bool rules = false;
  for (int k = 0; k < s.n_strategies; k++)
    if (s.strategies[k].type != LETF_ST_BENCHMARK && s.strategies[k].type != LETF_ST_BENCHMARK_TLT) rules = true;
  // 1: exactly a typed list; 2: a subset of the default sweep (routed); 3: dynamic slots
  const KernelEntry *best2 = nullptr, *best3 = nullptr;
  for (int tier = 1; tier <= 3 && !plan->cons; tier++) {
    if (tier == 3 && best2) {        // the cheapest routed list that holds the request
      bool r2 = false;
      route_request(s, best2->list, best2->ns, plan->slot_of, &r2);
      plan->cons = best2; plan->mask = best2->am & 0x3F; plan->routed = true;
      plan->n_slots = routed_slot_count(s, best2->list, plan->slot_of);
      break;
    }
    for (int i = 0; i < kernel_count() && !plan->cons; i++) {
      const KernelEntry* k = kernel_at(i);
      if (k->precision != LETF_PREC_F32 || k->market != kMarketFSeries || k->ns < s.n_strategies) continue;
      // (a frame with its own Cash_Ret column: the typed kernels do not read one - the routed kernel takes the whole list)
      if (tier == 1 && !series->cash_ret && k->list && !slot_list_routed(k->list) && (k->am & 0x3F) == mask && list_matches(k, e)) { plan->cons = k; plan->mask = mask; }
      if (tier == 2 && slot_list_routed(k->list) && !(mask & ~(k->am & 0x3F)) && !e->no_routed_split &&
          (!best2 || slot_list_count(k->list) < slot_list_count(best2->list))) {
        bool r2 = false;
        int slot_of[LETF_MAX_STRATEGIES];
        if (!route_request(s, k->list, k->ns, slot_of, &r2)) continue;
        best2 = k;
      }
      if (tier == 3 && !k->list && (k->am & 0x3F) == 0x3F && !(e->type_mask & ~k->tm) && !e->no_generic_split &&
          (!best3 || k->ns < best3->ns)) best3 = k;
    }
  }
  if (!plan->cons && best3) { plan->cons = best3; plan->mask = 0x3F; }
  plan->rules = rules;
  plan->nact = __builtin_popcount((unsigned)plan->mask);
  return plan->cons != nullptr;
}

template <int NSC>
static int launch_series_split(LetfEngine* e, const SplitPlan& plan, cudaStream_t st, const LetfSeries* series, const LetfOutputs* out) {
  const KernelEntry* kc = plan.cons;
  cudaFuncAttributes fac;
  int mbc = 0;
  CK(kc->attrs(&fac, &mbc));
  if (mbc < 1) return fail(LETF_ERR_CUDA, "rule kernel cannot be resident on this device");
  static thread_local KArgs<float, NSC, kAssets> ka_c;
  static thread_local SeriesPrepArgs pa;
  memset(&ka_c, 0, sizeof(ka_c)); memset(&pa, 0, sizeof(pa));
  if (plan.routed) {
    static thread_local LetfParams rp;
    rp = routed_params(e->params, kc->list, plan.slot_of, plan.n_slots, plan.mask);
    convert_params<float, NSC, kAssets>(rp, e->dev_dur, ka_c.p, (const DevBoot<float>*)(const void*)&e->boot32);
    route_slots(ka_c.p, e->params, plan.slot_of, plan.n_slots);
  } else {
    convert_params<float, NSC, kAssets>(e->params, e->dev_dur, ka_c.p, (const DevBoot<float>*)(const void*)&e->boot32);
  }
  const int T = series->sim_days, NF = plan.rules ? plan.nact + 2 : 0;
  const int64_t n = series->n_paths;
  // chunk of paths: as many as fit in 60 % of the free device memory (decided once per shape, see launch_split)
  const int64_t shape_key = ((int64_t)T << 32) ^ ((int64_t)NF << 24) ^ 1;
  int64_t cap = e->series_cap;
  if (e->series_key != shape_key || cap < 1) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const double per_path = (double)T * NF * 4;
    cap = per_path > 0 ? (int64_t)(0.6 * ((double)free_b + (double)e->rec.cap) / per_path) & ~(int64_t)127 : n;
    if (cap < 128) return fail(LETF_ERR_NOMEM, "materialised-path mode: the per-day records of 128 paths do not fit in device memory");
    e->series_key = shape_key; e->series_cap = cap;
  }
  if (e->force_series_chunk >= 128 && e->force_series_chunk < cap) cap = e->force_series_chunk & ~(int64_t)127;
  if (cap > n) cap = n;
  {
    cudaError_t ce = e->rec.reserve((size_t)cap * T * NF * 4);
    if (ce != cudaSuccess) { cudaGetLastError(); e->series_key = 0; return fail(LETF_ERR_NOMEM, "materialised-path records: " + std::string(cudaGetErrorString(ce))); }
  }
  pa.series = *series; pa.rec = NF ? (float*)e->rec.p : nullptr; pa.rec_ncol = cap; pa.T = T;
  if (series->cash_ret && plan.rules) {      // the frame's own Cash_Ret column (strategy.py:138 reads it per day)
    cudaError_t ce = e->rec_cash.reserve((size_t)cap * T * 4);
    if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "materialised-path cash column: " + std::string(cudaGetErrorString(ce))); }
    pa.rec_cash = (float*)e->rec_cash.p; ka_c.rec_cash = pa.rec_cash;
  }
  pa.field_mask = plan.mask; pa.asset_mask = ka_c.p.asset_mask; pa.bench_mask = ka_c.p.bench_mask; pa.bench_tlt = ka_c.p.bench_tlt;
  pa.n_strat = ka_c.p.n_strat; pa.W0 = (double)ka_c.p.W0;
  for (int q = 0; q < ka_c.p.n_strat && q < LETF_MAX_STRATEGIES; q++) {
    pa.type[q] = ka_c.p.strat[q].type; pa.asset[q] = ka_c.p.strat[q].asset; pa.route[q] = ka_c.p.strat[q].route;
  }
  ka_c.rng_mode = LETF_RNG_PHILOX; ka_c.rec = (float*)e->rec.p; ka_c.rec_ncol = cap; ka_c.log_trial = -1;
  ka_c.p.T = T; ka_c.p.buffer = 0; ka_c.p.random_start = 0; ka_c.series.sim_days = T;
  size_t scratch = (size_t)cap * T * NF * 4;
  int launches = 0;
  int64_t grid_last = 0;
  CK(cudaEventRecord(e->ev0, st));
  for (int64_t c0 = 0; c0 < n; c0 += cap) {
    const int64_t nc = n - c0 < cap ? n - c0 : cap;
    const LetfOutputs oc = out_columns(*out, c0);
    pa.out = oc; pa.col0 = c0; pa.n = nc;
    CK(launch_series_records(pa, st));
    launches++;
    if (!plan.rules) continue;
    const int64_t grid_c = full_wave_grid(nc, e->sm_count, mbc, kc->block);
    const int64_t nth = grid_c * kc->block;
    if ((int64_t)kRingP * nth >= ((int64_t)1 << 32)) return fail(LETF_ERR_INVALID, "grid too large for 32-bit ring offsets");
    ka_c.out = oc; ka_c.first_trial = c0; ka_c.n_trials = nc; ka_c.nthreads = nth; ka_c.series.n_paths = nc;
    if (kc->tm & kNeedRingP) { size_t b = (size_t)kRingP * nth * 4; CK(e->ringP.reserve(b)); ka_c.ringP = (float*)e->ringP.p; if (c0 == 0) scratch += b; }
    if (kc->tm & kNeedRingS) { size_t b = (size_t)kRingS * nth * 4; CK(e->ringS.reserve(b)); ka_c.ringS = (float*)e->ringS.p; if (c0 == 0) scratch += b; }
    if (kc->tm & kVolTypes) { size_t b = (size_t)kRingA * nth * 4; CK(e->ringA.reserve(b)); ka_c.ringA = (float*)e->ringA.p; if (c0 == 0) scratch += b; }
    CK(kc->launch(&ka_c, (int)grid_c, st));
    launches++; grid_last = grid_c;
  }
  CK(cudaEventRecord(e->ev1, st));
  e->timed = true;
  e->info.grid = (int)grid_last; e->info.block = kc->block; e->info.regs_per_thread = fac.numRegs;
  e->info.kernel_launches = launches; e->info.scratch_bytes = (int64_t)scratch; e->info.sm_count = e->sm_count;
  if (plan.rules) snprintf(e->info.kernel, sizeof(e->info.kernel), "series_records+%s", kc->name + 4);
  else snprintf(e->info.kernel, sizeof(e->info.kernel), "series_records");
  return LETF_OK;
}

static int launch_any(LetfEngine* e, int market, cudaStream_t st, uint64_t seed, int64_t first, int64_t n,
                      int rng_mode, const LetfTape* tape, const LetfOutputs* out, const LetfSeries* series,
                      const LogArgs& lg) {
  if (!e) return fail(LETF_ERR_INVALID, "null engine");
  if (!out || !out->final_wealth || !out->max_dd || !out->num_trades) return fail(LETF_ERR_INVALID, "outputs: final_wealth, max_dd and num_trades are required");
  if (n < 0 || first < 0) return fail(LETF_ERR_INVALID, "negative trial range");
  if (out->n_trials < n) return fail(LETF_ERR_INVALID, "outputs.n_trials smaller than n_trials");
  if (rng_mode < LETF_RNG_PHILOX || rng_mode > LETF_RNG_PHILOX_RECORD) return fail(LETF_ERR_INVALID, "bad rng_mode");
  if (rng_mode != LETF_RNG_PHILOX && market != kMarketSeries) {
    if (!tape) return fail(LETF_ERR_INVALID, "rng_mode needs a tape");
    const int G = e->params.sim_days + e->params.buffer_days;
    if (tape->gen_days != G || tape->sim_days != e->params.sim_days) return fail(LETF_ERR_INVALID, "tape day counts do not match the engine parameters");
    int64_t last = first + n - 1;
    int64_t need = n > 0 ? (e->params.antithetic_pairs ? (last >> 1) : last) + 1 : 0;
    if (tape->n_tapes < need) return fail(LETF_ERR_INVALID, "tape has fewer columns than the trial range needs");
    if (market == LETF_MARKET_BOOTSTRAP && (!tape->boot_len || !tape->boot_u || !tape->boot_t || tape->max_blocks < 1))
      return fail(LETF_ERR_INVALID, "bootstrap tape arrays missing");
    if (market == LETF_MARKET_BOOTSTRAP) {
      const LetfBootstrap& bb = e->params.boot;
      const bool empty = !bb.pool_a[0].n_blocks || !bb.pool_a[1].n_blocks || !bb.pool_b[0].n_blocks || !bb.pool_b[1].n_blocks;
      if (empty && (!tape->boot_syn || tape->boot_block_size != bb.block_size))
        return fail(LETF_ERR_INVALID, "a bootstrap pool is empty: the tape needs boot_syn with boot_block_size == block_size");
    }
  }
  if (n == 0) { e->info.kernel_launches = 0; e->info.kernel[0] = 0; return LETF_OK; }
  const bool want_journal = e->any_taxable && (out->journal_count || out->journal_volume);
  // (a caller-supplied frame needs no tape plumbing; only the journal does)
  const bool need_io = (market != kMarketSeries && rng_mode != LETF_RNG_PHILOX) || want_journal || lg.count != nullptr ||
                       lg.curve != nullptr || lg.market != nullptr;
  // a caller-supplied frame carries all six series; otherwise only the masked ones are simulated
  const int n_assets = market == kMarketSeries ? kAssets : __builtin_popcount((unsigned)e->params.asset_mask & 0x3Fu);
  // LETF_KERNEL=<a fused instantiation> pins the fused path; naming a producer / rule kernel picks among the split ones
  const bool pin_fused = e->pin_kernel.size() && e->pin_kernel.find("_prod_") == std::string::npos &&
                         e->pin_kernel.find("_rules_") == std::string::npos;
  const bool may_split = market == e->params.market_mode && !series && !pin_fused;
  if (may_split) {
    SplitPlan plan;
    if (find_split(e, rng_mode, need_io, first, &plan)) {
      CK(cudaSetDevice(e->device));
      if (plan.cons->ns == 8) return launch_split<8, 4>(e, plan, st, seed, first, n, out);
      if (plan.cons->ns == 32) return launch_split<32, 4>(e, plan, st, seed, first, n, out);
    }
  }
  if (market == kMarketSeries && series && !(e->pin_kernel.size() && e->pin_kernel.find("_rules_") == std::string::npos)) {
    SplitPlan plan;
    if (find_series_split(e, series, need_io, &plan)) {
      CK(cudaSetDevice(e->device));
      if (plan.cons->ns == 8) return launch_series_split<8>(e, plan, st, series, out);
      if (plan.cons->ns == 32) return launch_series_split<32>(e, plan, st, series, out);
    }
  }
  const KernelEntry* ke = select_kernel(e, market, need_io, e->params.n_strategies, e->type_mask, n_assets);
  // no typed pipeline and no antithetic-pair kernel for this list: the generic split pipeline before the fused builds
  if (may_split && !(ke && ke->fj) && !e->no_jpair) {
    SplitPlan plan;
    if (find_routed_split(e, rng_mode, need_io, first, &plan)) {
      CK(cudaSetDevice(e->device));
      if (plan.cons->ns == 8) return launch_split<8, 4>(e, plan, st, seed, first, n, out);
      return launch_split<32, 4>(e, plan, st, seed, first, n, out);
    }
    if (find_generic_split(e, rng_mode, need_io, first, &plan)) {
      CK(cudaSetDevice(e->device));
      if (plan.cons->ns == 8) return launch_split<8, kAssets>(e, plan, st, seed, first, n, out);
      return launch_split<32, kAssets>(e, plan, st, seed, first, n, out);
    }
  }
  if (!ke) return fail(LETF_ERR_INVALID, "no kernel instantiation for this (precision, market, strategies) combination");
  CK(cudaSetDevice(e->device));
#define GO(R, NS, NA) return launch_typed<R, NS, NA>(e, ke, st, seed, first, n, rng_mode, tape, out, series, lg)
  if (ke->precision == LETF_PREC_F64) {
    if (ke->ns == 32 && ke->na == 6) GO(double, 32, 6);
  } else {
    if (ke->ns == 4 && ke->na == 3) GO(float, 4, 3);
    if (ke->ns == 8 && ke->na == 6) GO(float, 8, 6);
    if (ke->ns == 32 && ke->na == 6) GO(float, 32, 6);
  }
#undef GO
  return fail(LETF_ERR_INVALID, "kernel registry holds an unsupported slot count");
}

// ---- 3x3 symmetric eigen-decomposition (cyclic Jacobi), for letf_nearest_psd3 -----------------
static void jacobi3(double a[3][3], double v[3][3], double w[3]) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
      if (fabs(a[p][q]) < 1e-300) continue;
      double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; k++) { double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
      for (int k = 0; k < 3; k++) { double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
      for (int k = 0; k < 3; k++) { double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < 3; i++) w[i] = a[i][i];
}

}  // namespace letf

// ================================================================================================
extern "C" {

int letf_abi_version(void) { return LETF_ABI_VERSION; }
const char* letf_last_error(void) { return g_err.c_str(); }

int letf_create(const LetfParams* params, int device, int precision, LetfEngine** out) {
  if (!params || !out) return fail(LETF_ERR_INVALID, "null argument");
  *out = nullptr;
  if (precision != LETF_PREC_F64 && precision != LETF_PREC_F32) return fail(LETF_ERR_INVALID, "bad precision");
  const char* err = validate_params(*params);
  if (err) return fail(LETF_ERR_INVALID, err);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(LETF_ERR_INVALID, "no such CUDA device");
  CK(cudaSetDevice(device));
  LetfEngine* e = new (std::nothrow) LetfEngine();
  if (!e) return fail(LETF_ERR_NOMEM, "host allocation failed");
  e->params = *params;
  e->device = device; e->precision = precision;
  memset(&e->info, 0, sizeof(e->info));
  int sms = 0;
  cudaError_t ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (ce != cudaSuccess) { delete e; return fail_cuda(ce, "cudaDeviceGetAttribute"); }
  e->sm_count = sms;
  e->info.sm_count = e->sm_count;
  for (int r = 0; r < 2; r++) {
    int nd = params->n_duration[r];
    if (nd > 0) {
      if (!params->duration_samples[r]) { delete e; return fail(LETF_ERR_INVALID, "n_duration > 0 with null duration_samples"); }
      e->host_dur[r].assign(params->duration_samples[r], params->duration_samples[r] + nd);
      ce = cudaMalloc((void**)&e->dev_dur[r], sizeof(int32_t) * nd);
      if (ce == cudaSuccess) ce = cudaMemcpy(e->dev_dur[r], e->host_dur[r].data(), sizeof(int32_t) * nd, cudaMemcpyHostToDevice);
      if (ce != cudaSuccess) { letf_destroy(e); return fail_cuda(ce, "duration samples upload"); }
      e->params.duration_samples[r] = e->host_dur[r].data();
    } else e->params.duration_samples[r] = nullptr;
  }
  memset(&e->boot64, 0, sizeof(e->boot64)); memset(&e->boot32, 0, sizeof(e->boot32));
  if (params->market_mode == LETF_MARKET_BOOTSTRAP) {
    // own copies of the caller's host tables, then one device blob: history [3][len], candidate rows, candidate returns
    LetfBootstrap& b = e->params.boot;
    const double* hsrc[3] = {params->boot.hist_spy, params->boot.hist_qqq, params->boot.hist_tlt};
    for (int c = 0; c < 3; c++) e->host_hist[c].assign(hsrc[c], hsrc[c] + b.hist_len);
    b.hist_spy = e->host_hist[0].data(); b.hist_qqq = e->host_hist[1].data(); b.hist_tlt = e->host_hist[2].data();
    for (int q = 0; q < 2; q++) for (int r = 0; r < 2; r++) {
      LetfBootstrapPool& pl = q ? b.pool_b[r] : b.pool_a[r];
      e->host_pstart[q][r].assign(pl.start, pl.start + pl.n_blocks);
      e->host_pret[q][r].assign(pl.block_ret, pl.block_ret + pl.n_blocks);
      pl.start = e->host_pstart[q][r].data(); pl.block_ret = e->host_pret[q][r].data();
    }
    auto upload = [&](auto& tables, auto& meta) -> cudaError_t {
      typedef typename std::remove_reference<decltype(tables.hist[0])>::type RR;
      const size_t hb = align256_(tables.hist.size() * sizeof(RR)), sb = align256_(tables.cstart.size() * sizeof(int)),
                   rb = tables.cret.size() * sizeof(RR);
      cudaError_t ce2 = cudaMalloc(&e->boot_blob, hb + sb + rb + 256);
      if (ce2 != cudaSuccess) return ce2;
      char* base = (char*)e->boot_blob;
      if ((ce2 = cudaMemcpy(base, tables.hist.data(), tables.hist.size() * sizeof(RR), cudaMemcpyHostToDevice)) != cudaSuccess) return ce2;
      if ((ce2 = cudaMemcpy(base + hb, tables.cstart.data(), tables.cstart.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess) return ce2;
      if ((ce2 = cudaMemcpy(base + hb + sb, tables.cret.data(), rb, cudaMemcpyHostToDevice)) != cudaSuccess) return ce2;
      meta = tables.meta;
      for (int c = 0; c < 3; c++) meta.hist[c] = (const RR*)base + (size_t)c * b.hist_len;
      meta.cstart = (const int*)(base + hb); meta.cret = (const RR*)(base + hb + sb);
      return cudaSuccess;
    };
    if (precision == LETF_PREC_F64) { BootTables<double> t; build_boot_tables<double>(e->params, t); ce = upload(t, e->boot64); }
    else { BootTables<float> t; build_boot_tables<float>(e->params, t); ce = upload(t, e->boot32); }
    if (ce != cudaSuccess) { letf_destroy(e); return fail_cuda(ce, "bootstrap tables upload"); }
  }
  e->type_mask = strategy_type_mask(*params);
  { const char* v = getenv("LETF_KERNEL"); e->pin_kernel = v ? v : ""; }
  { const char* v = getenv("LETF_NO_JPAIR"); e->no_jpair = v && *v && *v != '0'; }
  { const char* v = getenv("LETF_NO_SPLIT"); e->no_split = v && *v && *v != '0'; }
  { const char* v = getenv("LETF_NO_GENERIC_SPLIT"); e->no_generic_split = v && *v && *v != '0'; }
  { const char* v = getenv("LETF_NO_ROUTED_SPLIT"); e->no_routed_split = v && *v && *v != '0'; }
  { const char* v = getenv("LETF_TRACE_HOST"); e->trace_host = v && *v && *v != '0'; }
  { const char* v = getenv("LETF_SPLIT_WAVES"); e->force_split_waves = v ? atoi(v) : 0; }
  { const char* v = getenv("LETF_SERIES_CHUNK"); e->force_series_chunk = v ? atoll(v) : 0; }
  for (int k = 0; k < params->n_strategies; k++) if (params->strategies[k].taxable) e->any_taxable = true;
  ce = cudaEventCreate(&e->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev1);
  if (ce != cudaSuccess) { letf_destroy(e); return fail_cuda(ce, "cudaEventCreate"); }
  *out = e;
  return LETF_OK;
}

void letf_destroy(LetfEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (int r = 0; r < 2; r++) if (e->dev_dur[r]) cudaFree(e->dev_dur[r]);
  if (e->boot_blob) cudaFree(e->boot_blob);
  e->xscr.release(); e->ringP.release(); e->ringS.release(); e->ringA.release(); e->rec.release(); e->rec_cash.release(); e->spell.release();
  e->stage_tape.release(); e->stage_out.release(); e->stage_log.release(); e->pin_out.release();
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  delete e;
}

int letf_run(LetfEngine* e, void* stream, uint64_t seed, int64_t first_trial, int64_t n_trials, int rng_mode,
             const LetfTape* tape, const LetfOutputs* out) {
  if (!e) return fail(LETF_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  return launch_any(e, e->params.market_mode, (cudaStream_t)stream, seed, first_trial, n_trials, rng_mode, tape, out,
                    nullptr, LogArgs());
}

int letf_run_strategies(LetfEngine* e, void* stream, const LetfSeries* series, const LetfOutputs* out) {
  if (!e || !series) return fail(LETF_ERR_INVALID, "null argument");
  if (series->sim_days < 0 || series->n_paths < 0) return fail(LETF_ERR_INVALID, "bad series shape");
  if (!series->asset_ret[4]) return fail(LETF_ERR_INVALID, "series needs SPY_Ret (signals read SPY_Price)");
  std::lock_guard<std::mutex> lk(e->mu);
  return launch_any(e, kMarketSeries, (cudaStream_t)stream, 0, 0, series->n_paths, LETF_RNG_PHILOX, nullptr, out, series,
                    LogArgs());
}

int letf_summary_reset(LetfEngine* e, void* stream, const LetfSummary* sm) {
  if (!e || !sm || !sm->hist || !sm->moments || !sm->extrema || !sm->counters) return fail(LETF_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  CK(launch_summary_reset(*sm, e->params.n_strategies, (cudaStream_t)stream));
  e->info.aux_launches = 1;
  return LETF_OK;
}

int letf_reduce(LetfEngine* e, void* stream, const LetfOutputs* out, int64_t n_trials, const LetfSummary* sm) {
  if (!e || !out || !sm || !sm->hist || !sm->moments || !sm->extrema || !sm->counters) return fail(LETF_ERR_INVALID, "null argument");
  if (n_trials < 0 || n_trials > out->n_trials) return fail(LETF_ERR_INVALID, "bad n_trials");
  for (int q = 0; q < LETF_N_METRICS; q++) if (!(sm->hist_hi[q] > sm->hist_lo[q])) return fail(LETF_ERR_INVALID, "histogram range must have hi > lo");
  if (sm->bench_strategy >= e->params.n_strategies) return fail(LETF_ERR_INVALID, "bench_strategy out of range");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  CK(launch_reduce(*out, *sm, n_trials, e->params.n_strategies, e->params.initial_capital, e->params.sim_years,
                   e->sm_count, (cudaStream_t)stream));
  e->info.aux_launches = n_trials > 0 ? 1 : 0;
  return LETF_OK;
}

int64_t letf_summary_words(int n_strategies) {
  const int64_t S = n_strategies > 0 ? n_strategies : 1;
  return S * (LETF_N_METRICS * LETF_HIST_BINS + 4 + 2 * LETF_N_METRICS * 2);
}

int letf_summary_combine(LetfEngine* e, void* stream, const void* gathered, int world, int n_strategies, void* dst_packed) {
  if (!e || !gathered || !dst_packed) return fail(LETF_ERR_INVALID, "null argument");
  if (world < 1 || n_strategies < 1) return fail(LETF_ERR_INVALID, "bad world size / strategy count");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  CK(launch_summary_combine(gathered, world, n_strategies, dst_packed, (cudaStream_t)stream));
  e->info.aux_launches = 1;
  return LETF_OK;
}

int letf_launch_info(LetfEngine* e, LetfLaunchInfo* info) {
  if (!e || !info) return fail(LETF_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->timed) {
    CK(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->info.last_kernel_ms = ms;
  }
  *info = e->info;
  return LETF_OK;
}

// ---- host-buffer entry points ------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

constexpr int kTapeFields = 17;
struct TapeLayout { size_t off[kTapeFields]; size_t bytes[kTapeFields]; size_t total; };
static TapeLayout tape_layout(const LetfTape& t) {
  const size_t n = (size_t)t.n_tapes, G = (size_t)t.gen_days, T = (size_t)t.sim_days, K = (size_t)t.max_spells;
  const size_t B = t.boot_len ? (size_t)t.max_blocks : 0;
  const size_t b[kTapeFields] = {4 * n, 4 * K * n, 8 * G * 3 * n, 8 * G * n, 8 * G * n, 8 * G * n,
                                 8 * T * n, 8 * T * n, 8 * T * n, 8 * T * n, 8 * T * n, 8 * LETF_N_ASSETS * T * n,
                                 4 * B * n, 8 * B * 6 * n, B ? 8 * G * 3 * n : 0, t.start_regime ? 4 * n : 0,
                                 (t.boot_syn && B) ? 8 * B * 2 * (size_t)t.boot_block_size * n : 0};
  TapeLayout L; size_t o = 0;
  for (int i = 0; i < kTapeFields; i++) { L.off[i] = o; L.bytes[i] = b[i]; o += align256(b[i]); }
  L.total = o;
  return L;
}
static void** tape_field(LetfTape& t, int i) {
  switch (i) {
    case 0: return (void**)&t.start_offset; case 1: return (void**)&t.spells; case 2: return (void**)&t.x;
    case 3: return (void**)&t.chi2; case 4: return (void**)&t.vix_n; case 5: return (void**)&t.irx_n;
    case 6: return (void**)&t.liq_n; case 7: return (void**)&t.cred_n; case 8: return (void**)&t.jump_u;
    case 9: return (void**)&t.jump_t; case 10: return (void**)&t.sign; case 11: return (void**)&t.te_t;
    case 12: return (void**)&t.boot_len; case 13: return (void**)&t.boot_u; case 14: return (void**)&t.boot_t;
    case 15: return (void**)&t.start_regime; default: return (void**)&t.boot_syn;
  }
}

// stage a host tape on the device (host -> device copy of every array)
static int stage_tape(LetfEngine* e, const LetfTape* host, bool upload, LetfTape* dev) {
  *dev = *host;
  LetfTape h = *host;
  TapeLayout L = tape_layout(h);
  cudaError_t ce = e->stage_tape.reserve(L.total ? L.total : 256);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "tape staging: " + std::string(cudaGetErrorString(ce))); }
  for (int i = 0; i < kTapeFields; i++) {
    void* hp = *tape_field(h, i);
    char* dp = (char*)e->stage_tape.p + L.off[i];
    *tape_field(*dev, i) = (i >= 12 && !hp) ? nullptr : dp;       // the bootstrap arrays are optional
    if (L.bytes[i] == 0) continue;
    if (!hp) return fail(LETF_ERR_INVALID, "host tape has a null array");
    // record mode uploads too: entries the run does not consume keep the caller's values
    (void)upload;
    CK(cudaMemcpy(dp, hp, L.bytes[i], cudaMemcpyHostToDevice));
  }
  return LETF_OK;
}
static int unstage_tape(LetfEngine* e, const LetfTape* host) {
  LetfTape h = *host;
  TapeLayout L = tape_layout(h);
  for (int i = 0; i < kTapeFields; i++) {
    void* hp = *tape_field(h, i);
    if (L.bytes[i] && hp) CK(cudaMemcpy(hp, (char*)e->stage_tape.p + L.off[i], L.bytes[i], cudaMemcpyDeviceToHost));
  }
  return LETF_OK;
}

struct OutLayout { size_t off[9]; size_t bytes[9]; size_t total; };
static OutLayout out_layout(const LetfOutputs& o, int S, int64_t n) {
  const size_t s = (size_t)(S > 0 ? S : 1), m = (size_t)n;
  const size_t b[9] = {8 * s * m, 8 * s * m, 4 * s * m, o.journal_count ? 4 * s * m : 0, o.journal_volume ? 8 * s * m : 0,
                       o.start_offset ? 4 * m : 0, 4 * m /* flags always */, o.regime_days ? 4 * 2 * m : 0, o.jump_days ? 4 * m : 0};
  OutLayout L; size_t off = 0;
  for (int i = 0; i < 9; i++) { L.off[i] = off; L.bytes[i] = b[i]; off += align256(b[i]); }
  L.total = off;
  return L;
}

int letf_run_host(LetfEngine* e, uint64_t seed, int64_t first_trial, int64_t n_trials, int rng_mode,
                  const LetfTape* host_tape, const LetfOutputs* host_out) {
  if (!e || !host_out) return fail(LETF_ERR_INVALID, "null argument");
  if (n_trials < 0 || host_out->n_trials < n_trials) return fail(LETF_ERR_INVALID, "outputs.n_trials smaller than n_trials");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  LetfTape dtape; const LetfTape* tp = nullptr;
  if (rng_mode != LETF_RNG_PHILOX) {
    if (!host_tape) return fail(LETF_ERR_INVALID, "rng_mode needs a tape");
    int rc = stage_tape(e, host_tape, rng_mode == LETF_RNG_TAPE, &dtape);
    if (rc) return rc;
    tp = &dtape;
  }
  const int S = e->params.n_strategies;
  OutLayout L = out_layout(*host_out, S, n_trials);
  cudaError_t ce = e->stage_out.reserve(L.total ? L.total : 256);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "output staging: " + std::string(cudaGetErrorString(ce))); }
  char* base = (char*)e->stage_out.p;
  LetfOutputs dout;
  dout.n_trials = n_trials;     // device columns are packed [S][n_trials]
  dout.final_wealth = (double*)(base + L.off[0]); dout.max_dd = (double*)(base + L.off[1]);
  dout.num_trades = (int32_t*)(base + L.off[2]);
  dout.journal_count = L.bytes[3] ? (int32_t*)(base + L.off[3]) : nullptr;
  dout.journal_volume = L.bytes[4] ? (double*)(base + L.off[4]) : nullptr;
  dout.start_offset = L.bytes[5] ? (int32_t*)(base + L.off[5]) : nullptr;
  dout.flags = (uint32_t*)(base + L.off[6]);
  dout.regime_days = L.bytes[7] ? (int32_t*)(base + L.off[7]) : nullptr;
  dout.jump_days = L.bytes[8] ? (int32_t*)(base + L.off[8]) : nullptr;
  int rc = launch_any(e, e->params.market_mode, (cudaStream_t)0, seed, first_trial, n_trials, rng_mode, tp, &dout, nullptr, LogArgs());
  if (rc) return rc;
  if (n_trials == 0) return LETF_OK;
  // one device -> pinned copy of the whole staging block, then scatter device [S][n] -> host [S][N]
  // (row pitch = host capacity) with plain memcpy
  const size_t n = (size_t)n_trials, N = (size_t)host_out->n_trials;
  ce = e->pin_out.reserve(L.total ? L.total : 256);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "pinned staging: " + std::string(cudaGetErrorString(ce))); }
  CK(cudaMemcpyAsync(e->pin_out.p, base, L.total, cudaMemcpyDeviceToHost, (cudaStream_t)0));
  // while the kernel runs: fault in the caller's (typically freshly allocated, still unmapped) result pages, so the
  // scatter below is a plain copy instead of a few thousand first-touch page faults on the critical path
  {
    auto touch = [&](void* dst, int field, size_t elem, int nrows) {
      if (!dst || !L.bytes[field]) return;
      for (int r = 0; r < nrows; r++) {
        volatile char* q = (volatile char*)dst + (size_t)r * N * elem;
        const size_t len = n * elem;
        for (size_t o = 0; o < len; o += 4096) q[o] = 0;
        if (len) q[len - 1] = 0;
      }
    };
    touch(host_out->final_wealth, 0, 8, S); touch(host_out->max_dd, 1, 8, S); touch(host_out->num_trades, 2, 4, S);
    touch(host_out->journal_count, 3, 4, S); touch(host_out->journal_volume, 4, 8, S);
    touch(host_out->start_offset, 5, 4, 1); touch(host_out->flags, 6, 4, 1); touch(host_out->regime_days, 7, 4, 2);
    touch(host_out->jump_days, 8, 4, 1);
  }
  CK(cudaStreamSynchronize((cudaStream_t)0));
  const char* hb = (const char*)e->pin_out.p;
  // pinned -> the caller's (pageable) arrays.  One thread copies ~10 GB/s; a 1M-trial result block is 80 MB, so large rows
  // are split over a few host threads (the copy is the only part of a per-trial call that cannot overlap the kernel)
  auto big_copy = [](char* dst, const char* src, size_t bytes) {
    const size_t kMin = (size_t)4 << 20;
    unsigned nt = bytes >= 2 * kMin ? (unsigned)(bytes / kMin) : 1;
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned cap = hw >= 16 ? 8 : (hw >= 4 ? hw / 2 : 1);
    if (nt > cap) nt = cap;
    if (nt <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t part = ((bytes / nt) + 63) & ~(size_t)63;
    for (unsigned t = 1; t < nt; t++) {
      const size_t o = (size_t)t * part;
      if (o >= bytes) break;
      const size_t len = o + part < bytes ? part : bytes - o;
      th.emplace_back([=] { memcpy(dst + o, src + o, len); });
    }
    memcpy(dst, src, part < bytes ? part : bytes);
    for (auto& x : th) x.join();
  };
  auto rows = [&](void* dst, int field, size_t elem, int nrows) {
    if (!dst || !L.bytes[field]) return;
    if (N == n) { big_copy((char*)dst, hb + L.off[field], (size_t)nrows * n * elem); return; }     // same pitch: one block
    for (int r = 0; r < nrows; r++) big_copy((char*)dst + (size_t)r * N * elem, hb + L.off[field] + (size_t)r * n * elem, n * elem);
  };
  rows(host_out->final_wealth, 0, 8, S); rows(host_out->max_dd, 1, 8, S); rows(host_out->num_trades, 2, 4, S);
  rows(host_out->journal_count, 3, 4, S); rows(host_out->journal_volume, 4, 8, S);
  rows(host_out->start_offset, 5, 4, 1); rows(host_out->flags, 6, 4, 1); rows(host_out->regime_days, 7, 4, 2);
  rows(host_out->jump_days, 8, 4, 1);
  const uint32_t* fl = (const uint32_t*)(hb + L.off[6]);
  if (rng_mode == LETF_RNG_PHILOX_RECORD) { rc = unstage_tape(e, host_tape); if (rc) return rc; }
  int64_t bad = 0;
  for (size_t i = 0; i < n; i++) if (fl[i] & (LETF_FLAG_NONFINITE_RET | LETF_FLAG_BAD_VIX)) bad++;
  if (bad) {
    char msg[160];
    snprintf(msg, sizeof(msg), "%lld of %lld trials generated non-finite returns or VIX <= 0 (reference raises ValueError, engine.py:486-498)",
             (long long)bad, (long long)n_trials);
    return fail(LETF_ERR_NONFINITE, msg);
  }
  return LETF_OK;
}

int letf_trade_list_host(LetfEngine* e, uint64_t seed, int64_t trial, int strategy, int rng_mode,
                         const LetfTape* host_tape, int cap, int32_t* day, int32_t* asset, int32_t* action,
                         double* shares, double* price, int32_t* count) {
  if (!e || !count) return fail(LETF_ERR_INVALID, "null argument");
  if (strategy < 0 || strategy >= e->params.n_strategies) return fail(LETF_ERR_INVALID, "strategy out of range");
  if (cap < 0 || (cap > 0 && (!day || !asset || !action || !shares || !price))) return fail(LETF_ERR_INVALID, "bad log buffers");
  if (rng_mode == LETF_RNG_PHILOX_RECORD) return fail(LETF_ERR_INVALID, "trade list replay does not record");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  *count = 0;
  if (!e->params.strategies[strategy].taxable) return LETF_OK;   // engine.py:571: journal only for TAXABLE_IDS
  LetfTape dtape; const LetfTape* tp = nullptr;
  if (rng_mode == LETF_RNG_TAPE) {
    if (!host_tape) return fail(LETF_ERR_INVALID, "rng_mode needs a tape");
    int rc = stage_tape(e, host_tape, true, &dtape);
    if (rc) return rc;
    tp = &dtape;
  }
  const int S = e->params.n_strategies;
  const size_t c = (size_t)(cap > 0 ? cap : 1);
  size_t o_cnt = 0, o_day = 256, o_asset = o_day + align256(4 * c), o_act = o_asset + align256(4 * c),
         o_sh = o_act + align256(4 * c), o_pr = o_sh + align256(8 * c), o_out = o_pr + align256(8 * c);
  size_t total = o_out + align256(8 * S) * 2 + align256(4 * S) + 256;
  cudaError_t ce = e->stage_log.reserve(total);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "log staging"); }
  char* base = (char*)e->stage_log.p;
  CK(cudaMemset(base, 0, 256));
  LogArgs lg;
  lg.trial = trial; lg.strategy = strategy; lg.cap = cap;
  lg.count = (int32_t*)(base + o_cnt); lg.day = (int32_t*)(base + o_day); lg.asset = (int32_t*)(base + o_asset);
  lg.action = (int32_t*)(base + o_act); lg.shares = (double*)(base + o_sh); lg.price = (double*)(base + o_pr);
  LetfOutputs dout; memset(&dout, 0, sizeof(dout));
  dout.n_trials = 1;
  dout.final_wealth = (double*)(base + o_out); dout.max_dd = (double*)(base + o_out + align256(8 * S));
  dout.num_trades = (int32_t*)(base + o_out + 2 * align256(8 * S));
  dout.flags = (uint32_t*)(base + o_out + 2 * align256(8 * S) + align256(4 * S));
  int rc = launch_any(e, e->params.market_mode, (cudaStream_t)0, seed, trial, 1, rng_mode, tp, &dout, nullptr, lg);
  if (rc) return rc;
  int32_t nlog = 0;
  CK(cudaMemcpy(&nlog, lg.count, 4, cudaMemcpyDeviceToHost));
  *count = nlog;
  size_t m = (size_t)(nlog < cap ? nlog : cap);
  if (m) {
    CK(cudaMemcpy(day, lg.day, 4 * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(asset, lg.asset, 4 * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(action, lg.action, 4 * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(shares, lg.shares, 8 * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(price, lg.price, 8 * m, cudaMemcpyDeviceToHost));
  }
  return LETF_OK;
}

static int trade_lists_impl(LetfEngine* e, uint64_t seed, const int64_t* trials, const int32_t* strategies, int n, int strategy, int cap,
                            int32_t* day, int32_t* asset, int32_t* action, double* shares, double* price, int32_t* count);

int letf_trade_lists_host(LetfEngine* e, uint64_t seed, const int64_t* trials, int n, int strategy, int cap,
                          int32_t* day, int32_t* asset, int32_t* action, double* shares, double* price, int32_t* count) {
  if (!e) return fail(LETF_ERR_INVALID, "null argument");
  if (strategy < 0 || strategy >= e->params.n_strategies) return fail(LETF_ERR_INVALID, "strategy out of range");
  return trade_lists_impl(e, seed, trials, nullptr, n, strategy, cap, day, asset, action, shares, price, count);
}

int letf_trade_lists_multi_host(LetfEngine* e, uint64_t seed, const int64_t* trials, const int32_t* strategies, int n, int cap,
                                int32_t* day, int32_t* asset, int32_t* action, double* shares, double* price, int32_t* count) {
  if (!e || (n > 0 && !strategies)) return fail(LETF_ERR_INVALID, "null argument");
  for (int j = 0; j < n; j++)
    if (strategies[j] < 0 || strategies[j] >= e->params.n_strategies) return fail(LETF_ERR_INVALID, "strategy out of range");
  return trade_lists_impl(e, seed, trials, strategies, n, -1, cap, day, asset, action, shares, price, count);
}

static int trade_lists_impl(LetfEngine* e, uint64_t seed, const int64_t* trials, const int32_t* strategies, int n, int strategy, int cap,
                            int32_t* day, int32_t* asset, int32_t* action, double* shares, double* price, int32_t* count) {
  if (!e || !count || (n > 0 && !trials)) return fail(LETF_ERR_INVALID, "null argument");
  if (n < 0) return fail(LETF_ERR_INVALID, "negative trial count");
  if (cap < 0 || (cap > 0 && (!day || !asset || !action || !shares || !price))) return fail(LETF_ERR_INVALID, "bad log buffers");
  for (int j = 0; j < n; j++) { count[j] = 0; if (trials[j] < 0) return fail(LETF_ERR_INVALID, "negative trial id"); }
  if (n == 0) return LETF_OK;
  // engine.py:571: journal only for TAXABLE_IDS (a non-taxable strategy's rows stay empty)
  std::vector<int32_t> strat_host;
  if (strategies) {
    strat_host.assign(strategies, strategies + n);
    bool any = false;
    for (int j = 0; j < n; j++) { if (e->params.strategies[strat_host[j]].taxable) any = true; else strat_host[j] = -1; }
    if (!any) return LETF_OK;
  } else if (!e->params.strategies[strategy].taxable) return LETF_OK;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  const int S = e->params.n_strategies;
  const size_t c = (size_t)(cap > 0 ? cap : 1), m = (size_t)n;
  const size_t o_ids = 0, o_cnt = o_ids + align256(8 * m), o_day = o_cnt + align256(4 * m), o_asset = o_day + align256(4 * c * m),
               o_act = o_asset + align256(4 * c * m), o_sh = o_act + align256(4 * c * m), o_pr = o_sh + align256(8 * c * m),
               o_fw = o_pr + align256(8 * c * m), o_dd = o_fw + align256(8 * S * m), o_nt = o_dd + align256(8 * S * m),
               o_fl = o_nt + align256(4 * S * m), o_st = o_fl + align256(4 * m), total = o_st + align256(4 * m);
  cudaError_t ce = e->stage_log.reserve(total);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "log staging"); }
  char* base = (char*)e->stage_log.p;
  CK(cudaMemcpy(base + o_ids, trials, 8 * m, cudaMemcpyHostToDevice));
  CK(cudaMemset(base + o_cnt, 0, 4 * m));
  LogArgs lg;
  lg.trial_ids = (const int64_t*)(base + o_ids); lg.trial = -1; lg.strategy = strategy; lg.cap = cap;
  if (strategies) {
    CK(cudaMemcpy(base + o_st, strat_host.data(), 4 * m, cudaMemcpyHostToDevice));
    lg.strategies = (const int32_t*)(base + o_st);
  }
  lg.count = (int32_t*)(base + o_cnt); lg.day = (int32_t*)(base + o_day); lg.asset = (int32_t*)(base + o_asset);
  lg.action = (int32_t*)(base + o_act); lg.shares = (double*)(base + o_sh); lg.price = (double*)(base + o_pr);
  LetfOutputs dout; memset(&dout, 0, sizeof(dout));
  dout.n_trials = n;
  dout.final_wealth = (double*)(base + o_fw); dout.max_dd = (double*)(base + o_dd);
  dout.num_trades = (int32_t*)(base + o_nt); dout.flags = (uint32_t*)(base + o_fl);
  int rc = launch_any(e, e->params.market_mode, (cudaStream_t)0, seed, 0, n, LETF_RNG_PHILOX, nullptr, &dout, nullptr, lg);
  if (rc) return rc;
  CK(cudaMemcpy(count, lg.count, 4 * m, cudaMemcpyDeviceToHost));
  if (cap > 0) {
    // rows are [n][cap]; copy whole rows (the unused tail of a row is whatever the staging held)
    CK(cudaMemcpy(day, lg.day, 4 * c * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(asset, lg.asset, 4 * c * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(action, lg.action, 4 * c * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(shares, lg.shares, 8 * c * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(price, lg.price, 8 * c * m, cudaMemcpyDeviceToHost));
  }
  return LETF_OK;
}

// one path replayed with the journal log and / or the equity-curve log on; `series` (device pointers) selects the
// caller-supplied-frame mode.  Shared by letf_equity_curve_host and letf_run_strategy_curve_host.
static int replay_one(LetfEngine* e, int market, uint64_t seed, int64_t trial, int strategy, int rng_mode, const LetfTape* dtape,
                      const LetfSeries* dseries, int T, double* curve, int32_t* num_trades, int cap, int32_t* day, int32_t* asset,
                      int32_t* action, double* shares, double* price, int32_t* count) {
  const int S = e->params.n_strategies;
  const size_t c = (size_t)(cap > 0 ? cap : 1), t = (size_t)(T > 0 ? T : 1);
  const size_t o_cnt = 0, o_day = 256, o_asset = o_day + align256(4 * c), o_act = o_asset + align256(4 * c),
               o_sh = o_act + align256(4 * c), o_pr = o_sh + align256(8 * c), o_cv = o_pr + align256(8 * c),
               o_out = o_cv + align256(8 * t), total = o_out + align256(8 * S) * 2 + align256(4 * S) + 256;
  cudaError_t ce = e->stage_log.reserve(total);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "log staging"); }
  char* base = (char*)e->stage_log.p;
  CK(cudaMemset(base, 0, 256));
  CK(cudaMemset(base + o_cv, 0, 8 * t));
  LogArgs lg;
  lg.trial = trial; lg.strategy = strategy; lg.cap = cap;
  const bool want_log = count != nullptr && e->params.strategies[strategy].taxable;
  if (want_log) {
    lg.count = (int32_t*)(base + o_cnt); lg.day = (int32_t*)(base + o_day); lg.asset = (int32_t*)(base + o_asset);
    lg.action = (int32_t*)(base + o_act); lg.shares = (double*)(base + o_sh); lg.price = (double*)(base + o_pr);
  }
  if (curve) lg.curve = (double*)(base + o_cv);
  LetfOutputs dout; memset(&dout, 0, sizeof(dout));
  dout.n_trials = 1;
  dout.final_wealth = (double*)(base + o_out); dout.max_dd = (double*)(base + o_out + align256(8 * S));
  dout.num_trades = (int32_t*)(base + o_out + 2 * align256(8 * S));
  dout.flags = (uint32_t*)(base + o_out + 2 * align256(8 * S) + align256(4 * S));
  int rc = launch_any(e, market, (cudaStream_t)0, seed, trial, 1, rng_mode, dtape, &dout, dseries, lg);
  if (rc) return rc;
  if (curve && T > 0) CK(cudaMemcpy(curve, base + o_cv, 8 * (size_t)T, cudaMemcpyDeviceToHost));
  if (num_trades) CK(cudaMemcpy(num_trades, dout.num_trades + strategy, 4, cudaMemcpyDeviceToHost));
  if (count) *count = 0;
  if (want_log) {
    int32_t nlog = 0;
    CK(cudaMemcpy(&nlog, lg.count, 4, cudaMemcpyDeviceToHost));
    *count = nlog;
    const size_t m = (size_t)(nlog < cap ? nlog : cap);
    if (m) {
      CK(cudaMemcpy(day, lg.day, 4 * m, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(asset, lg.asset, 4 * m, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(action, lg.action, 4 * m, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(shares, lg.shares, 8 * m, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(price, lg.price, 8 * m, cudaMemcpyDeviceToHost));
    }
  }
  return LETF_OK;
}

int letf_equity_curve_host(LetfEngine* e, uint64_t seed, int64_t trial, int strategy, int rng_mode, const LetfTape* host_tape,
                           double* curve) {
  if (!e || !curve) return fail(LETF_ERR_INVALID, "null argument");
  if (strategy < 0 || strategy >= e->params.n_strategies) return fail(LETF_ERR_INVALID, "strategy out of range");
  if (rng_mode == LETF_RNG_PHILOX_RECORD) return fail(LETF_ERR_INVALID, "equity curve replay does not record");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  LetfTape dtape; const LetfTape* tp = nullptr;
  if (rng_mode == LETF_RNG_TAPE) {
    if (!host_tape) return fail(LETF_ERR_INVALID, "rng_mode needs a tape");
    int rc = stage_tape(e, host_tape, true, &dtape);
    if (rc) return rc;
    tp = &dtape;
  }
  return replay_one(e, e->params.market_mode, seed, trial, strategy, rng_mode, tp, nullptr, e->params.sim_days, curve, nullptr, 0,
                    nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int letf_market_returns_host(LetfEngine* e, uint64_t seed, int64_t trial, int rng_mode, const LetfTape* host_tape,
                             double* returns) {
  if (!e || !returns) return fail(LETF_ERR_INVALID, "null argument");
  if (trial < 0) return fail(LETF_ERR_INVALID, "negative trial id");
  if (rng_mode == LETF_RNG_PHILOX_RECORD) return fail(LETF_ERR_INVALID, "market replay does not record");
  if (e->params.n_strategies < 1) return fail(LETF_ERR_INVALID, "engine has no strategy slot");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  LetfTape dtape; const LetfTape* tp = nullptr;
  if (rng_mode == LETF_RNG_TAPE) {
    if (!host_tape) return fail(LETF_ERR_INVALID, "rng_mode needs a tape");
    int rc = stage_tape(e, host_tape, true, &dtape);
    if (rc) return rc;
    tp = &dtape;
  }
  const int S = e->params.n_strategies, T = e->params.sim_days;
  const size_t t3 = (size_t)(T > 0 ? T : 1) * 3;
  const size_t o_mk = 0, o_out = align256(8 * t3), total = o_out + align256(8 * S) * 2 + align256(4 * S) + 256;
  cudaError_t ce = e->stage_log.reserve(total);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "log staging"); }
  char* base = (char*)e->stage_log.p;
  CK(cudaMemset(base + o_mk, 0, 8 * t3));
  LogArgs lg;
  lg.trial = trial; lg.strategy = -1; lg.market = (double*)(base + o_mk);
  LetfOutputs dout; memset(&dout, 0, sizeof(dout));
  dout.n_trials = 1;
  dout.final_wealth = (double*)(base + o_out); dout.max_dd = (double*)(base + o_out + align256(8 * S));
  dout.num_trades = (int32_t*)(base + o_out + 2 * align256(8 * S));
  dout.flags = (uint32_t*)(base + o_out + 2 * align256(8 * S) + align256(4 * S));
  int rc = launch_any(e, e->params.market_mode, (cudaStream_t)0, seed, trial, 1, rng_mode, tp, &dout, nullptr, lg);
  if (rc) return rc;
  if (T > 0) CK(cudaMemcpy(returns, base + o_mk, 8 * (size_t)T * 3, cudaMemcpyDeviceToHost));
  return LETF_OK;
}

int letf_run_strategy_curve_host(LetfEngine* e, const LetfSeries* hs, int strategy, double* curve, int32_t* num_trades, int cap,
                                 int32_t* day, int32_t* asset, int32_t* action, double* shares, double* price, int32_t* count) {
  if (!e || !hs || !curve) return fail(LETF_ERR_INVALID, "null argument");
  if (hs->n_paths != 1 || hs->sim_days < 0) return fail(LETF_ERR_INVALID, "host frame: n_paths must be 1");
  if (strategy < 0 || strategy >= e->params.n_strategies) return fail(LETF_ERR_INVALID, "strategy out of range");
  if (!hs->asset_ret[4]) return fail(LETF_ERR_INVALID, "series needs SPY_Ret (signals read SPY_Price)");
  if (cap < 0 || (cap > 0 && (!day || !asset || !action || !shares || !price || !count))) return fail(LETF_ERR_INVALID, "bad log buffers");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  // upload the frame's columns into the tape staging buffer (unused in this mode)
  const size_t T = (size_t)hs->sim_days, col8 = align256(8 * (T ? T : 1)), col4 = align256(4 * (T ? T : 1));
  cudaError_t ce = e->stage_tape.reserve(col8 * (LETF_N_ASSETS + 3) + col4);
  if (ce != cudaSuccess) { cudaGetLastError(); return fail(LETF_ERR_NOMEM, "series staging"); }
  char* base = (char*)e->stage_tape.p;
  LetfSeries ds; memset(&ds, 0, sizeof(ds));
  ds.n_paths = 1; ds.sim_days = hs->sim_days;
  size_t o = 0;
  cudaError_t up_err = cudaSuccess;
  auto up = [&](const void* src, size_t bytes, size_t stride) -> const void* {
    if (!src) return nullptr;
    void* dst = base + o; o += stride;
    if (bytes) { cudaError_t ce2 = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice); if (ce2 != cudaSuccess && up_err == cudaSuccess) up_err = ce2; }
    return dst;
  };
  for (int a = 0; a < LETF_N_ASSETS; a++) ds.asset_ret[a] = (const double*)up(hs->asset_ret[a], 8 * T, col8);
  ds.tlt_ret = (const double*)up(hs->tlt_ret, 8 * T, col8);
  ds.vix = (const double*)up(hs->vix, 8 * T, col8);
  ds.regime = (const int32_t*)up(hs->regime, 4 * T, col4);
  ds.cash_ret = (const double*)up(hs->cash_ret, 8 * T, col8);
  CK(up_err);
  return replay_one(e, kMarketSeries, 0, 0, strategy, LETF_RNG_PHILOX, nullptr, &ds, hs->sim_days, curve, num_trades, cap, day, asset,
                    action, shares, price, count);
}

int letf_nearest_psd3(const double in[9], double out[9]) {
  if (!in || !out) return fail(LETF_ERR_INVALID, "null argument");
  double a[3][3], v[3][3], w[3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = 0.5 * (in[i * 3 + j] + in[j * 3 + i]);
  jacobi3(a, v, w);
  for (int i = 0; i < 3; i++) if (w[i] < 1e-8) w[i] = 1e-8;       // utils.py:45
  double m[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += v[i][k] * w[k] * v[j][k];
    m[i][j] = s;
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[i * 3 + j] = m[i][j] / (sqrt(m[i][i]) * sqrt(m[j][j]));
  return LETF_OK;
}

}  // extern "C"
