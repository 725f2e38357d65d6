// Materialised-path mode through the split strategy pipeline (run_strategy_fixed seam, letf/strategy.py:55).
//
// The caller's frame arrives as day-major fp64 series [day][path] (one array per ETF return column, TLT_Ret, VIX, an int32
// regime path).  The fused series build (Trial::run, MARKET = kMarketSeries) pulls them through registers next to 19 rules'
// state; here a bandwidth-bound pass turns them into the per-day fp32 records the rule kernels of the split pipeline
// already consume (letf_jfast.cuh EMIT layout: rec[(field * T + day) * ncol + column], the active return columns, VIX with
// the regime in its sign bit, SPY price) and does everything of a path that is not a rule: the buy&hold strategies
// (strategy.py:85-95), the validity flags and the regime day counts (engine.py:552-603).  One thread per path, coalesced
// 8-byte loads (every thread of a warp on the same day), four days of loads in flight per thread.
//
// Arithmetic is the fused series build's, operation for operation (returns rounded to fp32 first, fp64 price products,
// fp32 drawdown ratio), so the two routes agree bit for bit on everything this pass writes.
#pragma once
#include "letf_core.cuh"
#include "letf_jfast.cuh"      // f2u_ / u2f_

namespace letf {

struct SeriesPrepArgs {
  LetfSeries series;          // device pointers, pitch series.n_paths
  LetfOutputs out;            // view of the chunk's columns (pitch out.n_trials)
  float* rec;                 // records of this chunk, rec_ncol columns
  float* rec_cash;            // [day][rec_ncol] fp32 copy of the frame's Cash_Ret column, nullptr when the frame has none
  int64_t rec_ncol;
  int64_t col0, n;            // the chunk: paths [col0, col0 + n)
  int T;
  int field_mask;             // ETF return columns the rule kernel reads (its compile-time asset mask), in asset order
  int asset_mask;             // LetfParams::asset_mask (a buy&hold strategy on a series outside it reports W0, like the fused build)
  int bench_mask, bench_tlt;  // series tracked for a buy&hold strategy
  int n_strat;
  int type[LETF_MAX_STRATEGIES], asset[LETF_MAX_STRATEGIES], route[LETF_MAX_STRATEGIES];
  double W0;
};

// one path: j = column inside the chunk
LETF_HD void series_prep_path(const SeriesPrepArgs& a, const int64_t j) {
  const LetfSeries& s = a.series;
  const int T = a.T;
  const int64_t np = s.n_paths, col = a.col0 + j;
  double cp[kAssets], cp_tlt = 1.0;
  Drawdown<float> dd[kAssets], ddT;
#pragma unroll
  for (int k = 0; k < kAssets; k++) { cp[k] = 1.0; dd[k].init_open(); }
  ddT.init_open();
  float vmin = INFINITY, vmax = -INFINITY, chk = 0.0f;
  uint32_t flags = 0;
  int reg_days1 = 0;
  int nf = 0;
#pragma unroll
  for (int k = 0; k < kAssets; k++) nf += (a.field_mask >> k) & 1;
  const int64_t fstride = (int64_t)T * a.rec_ncol;
  // four days of loads in flight per thread (126 registers, 4 blocks per SM).  Measured on 100k paths x 2 520 days, 17.1 GB of
  // DRAM traffic: 3.68 ms; eight days (142 registers, 3 blocks per SM) 7.70 ms; two days under a 6-blocks-per-SM register cap
  // (one balanced wave, 108 B of spills) 4.17 ms
  constexpr int kAhead = 4;
  for (int i0 = 0; i0 < T; i0 += kAhead) {
    double er[kAhead][kAssets], vx[kAhead], tl[kAhead], ch[kAhead];
    int rg[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; d++) {
      const int64_t o = (int64_t)(i0 + d < T ? i0 + d : T - 1) * np + col;
#pragma unroll
      for (int k = 0; k < kAssets; k++) er[d][k] = s.asset_ret[k] ? s.asset_ret[k][o] : 0.0;
      vx[d] = s.vix ? s.vix[o] : 15.0;
      tl[d] = s.tlt_ret ? s.tlt_ret[o] : 0.0;
      rg[d] = s.regime ? s.regime[o] : 0;
      ch[d] = (s.cash_ret && a.rec_cash) ? s.cash_ret[o] : 0.0;
    }
#pragma unroll
    for (int d = 0; d < kAhead; d++) {
      const int i = i0 + d;
      if (i >= T) break;
      const float vix = (float)vx[d], tlt = (float)tl[d];
      const int reg = rg[d];
      chk = fmaf(tlt, 0.0f, chk);
      if (!isfinite_(vix) || vix <= 0.0f) flags |= LETF_FLAG_BAD_VIX;
      reg_days1 += reg;
      vmin = fminf(vmin, vix); vmax = fmaxf(vmax, vix);
      float* const row = a.rec ? a.rec + (int64_t)i * a.rec_ncol + j : nullptr;
      int f = 0;
#pragma unroll
      for (int k = 0; k < kAssets; k++) {
        const float r = (float)er[d][k];
        if (s.asset_ret[k]) {
          cp[k] = ::fma(cp[k], (double)r, cp[k]);
          if ((a.bench_mask >> k) & 1) dd[k].step_ret(r);
        }
        if (row && ((a.field_mask >> k) & 1)) { row[(int64_t)f * fstride] = s.asset_ret[k] ? r : 0.0f; f++; }
      }
      if (a.bench_tlt) { cp_tlt = ::fma(cp_tlt, (double)tlt, cp_tlt); ddT.step_ret(tlt); }
      if (a.rec_cash) a.rec_cash[(int64_t)i * a.rec_ncol + j] = (float)ch[d];
      if (row) {
        row[(int64_t)nf * fstride] = u2f_(f2u_(vix) | ((uint32_t)(reg != 0) << 31));
        row[(int64_t)(nf + 1) * fstride] = (float)(cp[4] * 100.0);
      }
    }
  }
  if (chk != 0.0f) flags |= LETF_FLAG_NONFINITE_RET;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < kAssets; k++) if (s.asset_ret[k] && T > 0 && !(cp[k] > 0.0 && isfinite_(cp[k]))) bad = true;
  if (a.bench_tlt && T > 0 && !(cp_tlt > 0.0 && isfinite_(cp_tlt))) bad = true;
  if (T == 0 || bad || !(vmin >= 5.0f) || !(vmax <= 120.0f)) flags |= LETF_FLAG_REPAIRED;
  const int64_t N = a.out.n_trials;
  for (int q = 0; q < a.n_strat; q++) {
    const int ty = a.type[q];
    if (ty != LETF_ST_BENCHMARK && ty != LETF_ST_BENCHMARK_TLT) continue;
    if (a.route[q] < 0) continue;
    const int64_t row = a.route[q] > 0 ? a.route[q] - 1 : q;
    double c = 1.0; Drawdown<float> d = ddT;
    if (ty == LETF_ST_BENCHMARK_TLT) c = cp_tlt;
    else {
#pragma unroll
      for (int k = 0; k < kAssets; k++) if (((a.asset_mask >> k) & 1) && k == a.asset[q]) { c = cp[k]; d = dd[k]; }
    }
    a.out.final_wealth[row * N + j] = T > 0 ? (double)(float)a.W0 * c : 0.0;
    a.out.max_dd[row * N + j] = T > 1 ? d.value() : 0.0;
    a.out.num_trades[row * N + j] = 0;
    if (a.out.journal_count) a.out.journal_count[row * N + j] = 0;
    if (a.out.journal_volume) a.out.journal_volume[row * N + j] = 0.0;
  }
  if (a.out.start_offset) a.out.start_offset[j] = 0;
  if (a.out.flags) a.out.flags[j] = flags;
  if (a.out.regime_days) { a.out.regime_days[j] = T - reg_days1; a.out.regime_days[N + j] = reg_days1; }
  if (a.out.jump_days) a.out.jump_days[j] = 0;
}

}  // namespace letf
