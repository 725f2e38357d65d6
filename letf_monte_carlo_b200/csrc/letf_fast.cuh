// letf_fast.cuh — the straight-line day step of the lean buy&hold builds, U trials per thread.
//
// Same computation as Trial::run_t<FAST> (letf_core.cuh) for the shapes DevParams::fast_layers admits
// (iid-normal market, every compacted slot active, t(5) tracking error, no tanh cap, no funding model, no
// TLT series, buy&hold only): reference layers engine.py:360-420, statistics engine.py:578-590.
// It exists for instruction-level parallelism: BASELINE configs[1] has 100 000 trials, i.e. only ~5 warps
// per scheduler, and one trial's day step is a handful of long dependent chains (10 Philox rounds, then
// lg2 -> rsqrt -> FMA chains per series).  Interleaving U = 2 independent trials in one thread doubles the
// independent chains the scheduler sees while the grid still fits in one wave (ceil(n / U) threads).
// Philox counters, draw order and arithmetic are those of run_t<FAST>, so results agree to fp32 contraction.
#pragma once
#include "letf_core.cuh"

namespace letf {

// K10: bitmask of the slots whose tracking-error innovations are t(10) (the others are t(5)); NACT: active (compacted)
// slots; TE = false: every active series has tracking_error_base == 0 (the frictionless vol-drag shape of
// validation.py:33-57, BASELINE config 1) - no tracking-error block is drawn at all
template <int NS, int NA, int U, int K10 = 0, int NACT = NA, bool TE = true>
struct FastLayers {
  typedef float R;
  typedef Mth<float> M;

  // trial[u] / col[u]: global trial id and output column; live[u]: false for the padding lane of an odd tail
  static LETF_HD void run(const KArgs<R, NS, NA>& ka, const int64_t (&trial)[U],
                                             const int64_t (&col)[U], const bool (&live)[U]) {
    const DevParams<R, NS, NA>& p = ka.p;
    const int T = p.T;
    RngKey key[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t sim_id = p.antithetic ? (trial[u] & ~(int64_t)1) : trial[u];
      key[u].k0 = (uint32_t)ka.seed; key[u].k1 = (uint32_t)(ka.seed >> 32);
      key[u].t0 = (uint32_t)sim_id; key[u].t1 = (uint32_t)((uint64_t)sim_id >> 32);
    }
    R te_prev[U][NACT], bias[U][NACT], q[U][NACT], qmin[U][NACT], chk[U];
    double cp[U][NACT];
#pragma unroll
    for (int u = 0; u < U; u++) {
      chk[u] = 0.0f;
#pragma unroll
      for (int a = 0; a < NACT; a++) { te_prev[u][a] = 0.0f; bias[u][a] = 0.0f; q[u][a] = INFINITY; qmin[u][a] = 1.0f; cp[u][a] = 1.0; }
    }
    const R vix = p.const_vix;
    const R vrmult = clipx(M::sqrt_(vix * 0.05f), 0.5f, 1.5f) * 1.0f;      // regime 0 (engine.py:178-181)
    uint32_t streams[NACT];
#pragma unroll
    for (int a = 0; a < NACT; a++) streams[a] = STREAM_TE0 + (uint32_t)p.act_id[a];
    const R alpha = (R)(2.0 / 127.0);

    // one simulated day of all U trials; (wa, wb) = the two market words of the day
    auto day = [&](const int i, const uint32_t (&wa)[U], const uint32_t (&wb)[U]) {
      U4 blk[U][NACT];
      if (TE) {
#pragma unroll
        for (int u = 0; u < U; u++) rng_blocks<NACT>(key[u], (uint32_t)i, streams, blk[u]);
      }
      const R vrlive = i >= 1 ? vrmult : 0.0f;          // no tracking error on day 0 (engine.py:184)
      R spy[U], qqq[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        R z0, z1;
        M::box_muller(wa[u], wb[u], z0, z1);
        spy[u] = p.n_mu[0] + p.n_ch[0] * z0;
        qqq[u] = p.n_mu[1] + (p.n_ch[1] * z0 + p.n_ch[2] * z1);
        chk[u] = fmaf(spy[u], 0.0f, chk[u]); chk[u] = fmaf(qqq[u], 0.0f, chk[u]);
      }
#pragma unroll
      for (int a = 0; a < NACT; a++) {
        const DevAsset<R>& ap = p.act[a];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const R un = fmaf(ap.w_spy, spy[u], ap.w_qqq * qqq[u]);
          const R lev = ap.lev * un - ap.carry_const;
          R er;
          if (TE) {
            const R innov = ((K10 >> a) & 1) ? student_t10<R>(blk[u][a]) : student_t5<R>(blk[u][a]);
            const R ds = un < 0.0f ? ap.te_down : 0.98f;
            const R mm = (1.0f + 0.3f * fabsf(un)) * ds;
            const R te_raw = ap.te_rho * te_prev[u][a] + (innov * vrlive) * (ap.te_scale * mm);
            te_prev[u][a] = te_raw;
            bias[u][a] = alpha * te_raw + (1.0f - alpha) * bias[u][a];
            er = clipfast(lev + (te_raw - bias[u][a]), -0.999f, 10.0f);
          } else {
            er = clipfast(lev, -0.999f, 10.0f);       // te_scale == 0: the AR(1) residual and its EMA bias are identically 0
          }
          cp[u][a] = ::fma(cp[u][a], (double)er, cp[u][a]);
          q[u][a] = fminf(1.0f, fmaf(q[u][a], er, q[u][a]));
          qmin[u][a] = fminf(qmin[u][a], q[u][a]);
        }
      }
    };

    // days 2j and 2j+1 share the market block (j, STREAM_MKT_PAIR): words x,y then z,w (letf_core.cuh market_day)
    int i = 0;
    for (; i + 1 < T; i += 2) {
      uint32_t w0[U], w1[U], w2[U], w3[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const U4 bp = rng_block(key[u], (uint32_t)i >> 1, STREAM_MKT_PAIR);
        w0[u] = bp.x; w1[u] = bp.y; w2[u] = bp.z; w3[u] = bp.w;
      }
      day(i, w0, w1);
      day(i + 1, w2, w3);
    }
    if (i < T) {
      uint32_t w0[U], w1[U];
#pragma unroll
      for (int u = 0; u < U; u++) { const U4 bp = rng_block(key[u], (uint32_t)i >> 1, STREAM_MKT_PAIR); w0[u] = bp.x; w1[u] = bp.y; }
      day(i, w0, w1);
    }

    // per-trial outputs (engine.py:578-603); validity as in run_t: constant VIX, sticky non-finite / exhausted products
    const int64_t N = ka.out.n_trials;
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (!live[u]) continue;
      uint32_t flags = 0;
      bool bad = false;
      if (T > 0 && (!isfinite_(vix) || vix <= 0.0f)) flags |= LETF_FLAG_BAD_VIX;
      if (chk[u] != 0.0f) flags |= LETF_FLAG_NONFINITE_RET;
#pragma unroll
      for (int a = 0; a < NACT; a++) if (T > 0 && !(cp[u][a] > 0.0 && isfinite_(cp[u][a]))) bad = true;
      if (T == 0 || bad || !(vix >= 5.0f) || !(vix <= 120.0f)) flags |= LETF_FLAG_REPAIRED;
      for (int s = 0; s < NS; s++) {
        if (s >= p.n_strat) break;
        double c = 1.0; R qm = 1.0f;
#pragma unroll
        for (int a = 0; a < NACT; a++) if (p.act_id[a] == p.strat[s].asset) { c = cp[u][a]; qm = qmin[u][a]; }
        ka.out.final_wealth[(int64_t)s * N + col[u]] = T > 0 ? (double)p.W0 * c : 0.0;
        ka.out.max_dd[(int64_t)s * N + col[u]] = T > 1 ? 1.0 - (double)qm : 0.0;
        ka.out.num_trades[(int64_t)s * N + col[u]] = 0;
        if (ka.out.journal_count) ka.out.journal_count[(int64_t)s * N + col[u]] = 0;
        if (ka.out.journal_volume) ka.out.journal_volume[(int64_t)s * N + col[u]] = 0.0;
      }
      if (ka.out.start_offset) ka.out.start_offset[col[u]] = 0;
      if (ka.out.flags) ka.out.flags[col[u]] = flags;
      if (ka.out.regime_days) { ka.out.regime_days[col[u]] = T; ka.out.regime_days[N + col[u]] = 0; }
      if (ka.out.jump_days) ka.out.jump_days[col[u]] = 0;
    }
  }
};

// DevParams::fast_layers = (1 + mask of the t(10) slots) | n_act << 8 | (no tracking error) << 12.  Shapes with a
// compiled-in straight-line loop: all NA slots t(5) (BASELINE config 2), "last slot t(10)" (config 5: UPRO / SSO / SPY),
// and one frictionless series (config 1).  Any other admitted shape runs the general loop.
LETF_HD constexpr int fast_layers_code(int nact, int k10, bool no_te) { return (1 + k10) | (nact << 8) | ((no_te ? 1 : 0) << 12); }
template <int NA> LETF_HD constexpr bool fast_pattern_compiled(int code) {
  return code == fast_layers_code(NA, 0, false) || code == fast_layers_code(NA, 1 << (NA - 1), false) ||
         code == fast_layers_code(1, 0, true) || code == fast_layers_code(1, 1, true);
}

template <int NS, int NA, int U>
LETF_HD void fast_layers_run(const KArgs<float, NS, NA>& ka, const int64_t (&trial)[U], const int64_t (&col)[U],
                             const bool (&live)[U]) {
  const int code = ka.p.fast_layers;
  if (code == fast_layers_code(NA, 0, false)) FastLayers<NS, NA, U, 0>::run(ka, trial, col, live);
  else if (code == fast_layers_code(NA, 1 << (NA - 1), false)) FastLayers<NS, NA, U, 1 << (NA - 1)>::run(ka, trial, col, live);
  else FastLayers<NS, NA, U, 0, 1, false>::run(ka, trial, col, live);
}

}  // namespace letf
