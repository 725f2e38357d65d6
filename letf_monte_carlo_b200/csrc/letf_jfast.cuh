// letf_jfast.cuh — the institutional (joint-t) market for lean buy&hold shapes, one ANTITHETIC PAIR per thread.
//
// Same computation as Trial::run (letf_core.cuh) for the shapes DevParams::fast_joint admits: joint-t
// GARCH/DCC market with moment matching (calibration.py:480-590), VIX / IRX (engine.py:100-123), latent stress +
// jump overlay (calibration.py:787-830, engine.py:502-512), state-linked financing (engine.py:320-349,
// calibration.py:833-857), AR(1) Student-t tracking error with default dof and no tanh cap (engine.py:157-203),
// layer composition (engine.py:360-420), buy&hold statistics (engine.py:578-590), Philox draws only.
//
// Why a pair per thread: the reference seeds both twins of an antithetic pair identically (mc_runner.py:145-153,
// engine.py:436) and only negates the correlated normal vector (calibration.py:563), so the twins share the regime
// path, the start offset, every uniform / normal / chi-square / Student-t variate, the IRX path and the liquidity /
// credit channels.  One thread therefore draws each Philox block once and turns it into variates once for both
// twins (the generic build does all of it twice), keeps only the genuinely path-dependent state twice (GARCH h, DCC
// Q, prev_z, VIX, tracking-error AR(1), equity products, drawdowns), and the two independent dependency chains give
// the scheduler instruction-level parallelism at 3 warps per scheduler.  Philox counters, draw order and arithmetic
// are those of Trial::run, so results agree with the generic build to fp32 reassociation.
//
// HBM: moment matching needs the column means of all G generated days before anything downstream, so pass 1 writes
// one 32-byte record per generated day and PAIR (thread-major: each thread streams its own column; one full sector,
// one 256-bit store): the three raw returns of both twins, the day's VIX normal, and the day's risk-free rate with the
// regime in its sign bit.  Pass 2 (and the VIX burn-in before the random start) reads it back with one 256-bit load
// per day, requested a day ahead.  16 B written + 16 B read per generated path-day.
#pragma once
#include "letf_core.cuh"

namespace letf {

struct alignas(32) JF8 { float a0, a1, a2, a3, b0, b1, b2, b3; };

LETF_HD JF8 jf_load(const JF8* p) {
#if defined(__CUDA_ARCH__)
  JF8 v;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.a0), "=f"(v.a1), "=f"(v.a2), "=f"(v.a3), "=f"(v.b0), "=f"(v.b1), "=f"(v.b2), "=f"(v.b3)
               : "l"(p));
  return v;
#else
  return *p;
#endif
}
LETF_HD void jf_store(JF8* p, const JF8& v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(v.a0), "f"(v.a1), "f"(v.a2), "f"(v.a3), "f"(v.b0), "f"(v.b1), "f"(v.b2), "f"(v.b3)
               : "memory");
#else
  *p = v;
#endif
}

LETF_HD uint32_t f2u_(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
LETF_HD float u2f_(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

// dwell time of spell `ks` (engine.py:255-302; same draw as Trial::draw_spell).  Out of line and by value: it runs
// once per regime spell, and nothing of the caller's state has to live in memory for it.
static LETF_HDN int jf_spell(uint32_t k0, uint32_t k1, uint32_t t0, uint32_t t1, uint32_t ks, float log_pstay, int n_dur,
                             const int32_t* dur, int G) {
  typedef Mth<float> M;
  RngKey key; key.k0 = k0; key.k1 = k1; key.t0 = t0; key.t1 = t1;
  const U4 b = rng_block(key, ks, STREAM_SPELL);
  int spell;
  if (n_dur > 0) spell = dur[mulhi32(b.y, (uint32_t)n_dur)];
  else {
    const float q = M::div_(M::log_(M::u01(b.x)), log_pstay);
    spell = q < (float)G ? (int)q + 1 : G;
  }
  return spell < 1 ? 1 : spell;
}

// the jump size of day g (calibration.py:828): |t(5)|, shared by both twins
static LETF_HDN float jf_jump_t5(uint32_t k0, uint32_t k1, uint32_t t0, uint32_t t1, uint32_t g) {
  RngKey key; key.k0 = k0; key.k1 = k1; key.t0 = t0; key.t1 = t1;
  return Mth<float>::abs_(student_t5<float>(rng_block(key, g, STREAM_JUMP)));
}

// NACT: active ETF series (compacted slots 0 .. NACT-1); K10: bitmask of the slots whose tracking-error innovations are
// t(10) (the others t(5)).
// EMIT: producer half of the split strategy pipeline (letf_api.cu launch_split).  Besides the buy&hold statistics the
// thread writes, per simulated day and twin, the record the rule kernel consumes - the NACT layer returns, the VIX with
// the regime in its sign bit, and the SPY price (last active slot) - day-major and field-major,
// rec[(field * T + day) * ncol + column], so that both its own float2 stores (twins are adjacent columns) and the rule
// kernel's loads (one thread per column, all on the same day) are coalesced.  Only benchmark-type strategy slots are
// written by the producer; the rule kernel (Trial::run, MARKET == kMarketFSeries) writes the others.
template <int NS, int NA, int NACT, int K10, bool EMIT = false>
struct JointPair {
  typedef float R;
  typedef Mth<float> M;
  static_assert(NACT >= 1 && NACT <= NA, "active slots");
  static_assert(NA < kAssets || NACT == kAssets, "uncompacted build: slot j is asset j");
  // uncompacted build (NA == 6, the producer of the generic split pipeline): slot j is asset j, series outside the asset mask
  // are skipped at run time (their record fields are written as zeros); SPY sits in slot 4, otherwise in the last active slot
  static constexpr bool ALLSLOTS = NA == kAssets;
  static constexpr int kSpySlot = ALLSLOTS ? 4 : NACT - 1;
  // K10 bit 7: tracking-error dof and tanh cap read at run time (calibrated tracking_residual_model, engine.py:185, :198-203)
  static constexpr bool TEDF = (K10 & 0x80) != 0;

  // pj: global pair index (trials 2 pj and 2 pj + 1); lane: this thread's scratch column
  // fjs / jts: the two regimes' per-day constants and joint-model constants staged in shared memory by the kernel
  // (letf_kernels.cuh).  A lane indexes them with ITS regime: from the constant bank that is a divergent LDC, replayed per
  // distinct address; from shared memory one LDS.  nullptr (host emulation): read them from the parameter block.
  static LETF_HD void run(const KArgs<R, NS, NA>& ka, const int64_t pj, const int64_t lane,
                          const DevFJ<R>* fjs = nullptr, const DevJoint<R>* jts = nullptr) {
    const DevParams<R, NS, NA>& p = ka.p;
    if (!fjs) fjs = p.fj;
    if (!jts) jts = p.joint;
    const int T = p.T, G = p.T + p.buffer;
    RngKey key;
    key.k0 = (uint32_t)ka.seed; key.k1 = (uint32_t)(ka.seed >> 32);
    { const int64_t sim_id = pj * 2; key.t0 = (uint32_t)sim_id; key.t1 = (uint32_t)((uint64_t)sim_id >> 32); }

    // ---- start offset / start regime (random_start.py:33-57, :169-199), shared by the twins
    int so = 0, cur = 0;
    if (p.random_start || p.start_p1 > 0.0f) {
      const U4 b = rng_block(key, 0, STREAM_START);
      if (p.random_start) so = p.min_offset + (int)mulhi32(b.x, (uint32_t)p.offset_range);
      if (p.start_p1 > 0.0f) cur = M::u01(b.y) < p.start_p1 ? 1 : 0;
    }

    JF8* const xs = (JF8*)ka.xscr + lane * (int64_t)G;

    // ---- regime path (engine.py:255-302), once per pair, before anything else: the dwell-time draws are a Philox block each
    // and happen on ~2 % of the days of a lane, i.e. on ~half of the days of a WARP - inside the day loop the whole warp
    // would step through the draw on each of those days.  Here every lane walks its own ~G/50 spells in a compact loop and
    // leaves (last day of the spell + 1) << 1 | regime in its scratch row; the day loop only compares against the next entry.
    uint32_t* const sw = ka.spell_scr + lane * (int64_t)(G + 1);
    {
      int day = 0;
      uint32_t ks = 0;
      while (day < G) {
        if (ks > 0 && p.flip[cur]) cur ^= 1;
        day += jf_spell(key.k0, key.k1, key.t0, key.t1, ks, p.log_pstay[cur], p.n_dur[cur], p.dur[cur], G);
        sw[ks++] = ((uint32_t)(day < G ? day : G) << 1) | (uint32_t)cur;
      }
    }

    // ---- pass 1: raw joint returns of both twins -> scratch, column means (calibration.py:508-588)
    float shift[2][3];
    {
      uint32_t ks = 0, nsw = 0;                 // nsw: first day of the next spell
      float h[2][3], q[2][6], pz[2][3], sum[2][3], cmp[2][3], part[2][3];
#pragma unroll
      for (int t = 0; t < 2; t++) {
#pragma unroll
        for (int a = 0; a < 3; a++) { h[t][a] = 0.0f; pz[t][a] = 0.0f; sum[t][a] = 0.0f; cmp[t][a] = 0.0f; part[t][a] = 0.0f; }
#pragma unroll
        for (int a = 0; a < 6; a++) q[t][a] = 0.0f;
      }
      // per-regime constants held in registers between regime switches
      float mu[3] = {0, 0, 0}, c0[3] = {0, 0, 0}, c1[6] = {0, 0, 0, 0, 0, 0};
      float alpha = 0, beta = 0, dcc_a = 0, dcc_b = 0, sqrt_nu = 0, kodd = 0, irx_base = 0, nu = 0;
      int kk = 0, jreg = -1;
      const uint32_t mstreams[2] = {STREAM_MKT0, STREAM_MKT1};
      int reg = 0;
      for (int g = 0; g < G; g++) {
        if ((uint32_t)g == nsw) { const uint32_t e = sw[ks++]; reg = (int)(e & 1u); nsw = e >> 1; }
        U4 w[2];
        rng_blocks<2>(key, (uint32_t)g, mstreams, w);
        float z0, z1, z2, nchi, vn, in;
        M::box_muller(w[0].x, w[0].y, z0, z1);
        M::box_muller(w[0].z, w[0].w, z2, nchi);
        M::box_muller(w[1].x, w[1].y, vn, in);
        if (reg != jreg) {
          // regime switch (or day 0): reload the constants; GARCH / DCC state resets on day 0 and otherwise blends
          // 10 % toward the new regime's long-run values (calibration.py:532-541)
          const DevJoint<R>& jp = jts[reg];
#pragma unroll
          for (int a = 0; a < 3; a++) { mu[a] = jp.mu[a]; c0[a] = jp.c0[a]; }
#pragma unroll
          for (int a = 0; a < 6; a++) c1[a] = jp.c1[a];
          alpha = jp.alpha; beta = jp.beta; dcc_a = jp.dcc_a; dcc_b = jp.dcc_b; sqrt_nu = jp.sqrt_nu;
          kk = jp.nu_fast_k; kodd = (kk & 1) ? 1.0f : 0.0f; nu = jp.nu;
          irx_base = p.irx_base[reg];
#pragma unroll
          for (int t = 0; t < 2; t++) {
#pragma unroll
            for (int a = 0; a < 3; a++) h[t][a] = g == 0 ? jp.lv[a] : 0.10f * jp.lv[a] + (1.0f - 0.10f) * h[t][a];
#pragma unroll
            for (int a = 0; a < 6; a++) q[t][a] = g == 0 ? jp.rb[a] : 0.10f * jp.rb[a] + (1.0f - 0.10f) * q[t][a];
          }
          jreg = reg;
        } else {
#pragma unroll
          for (int t = 0; t < 2; t++) {
#pragma unroll
            for (int a = 0; a < 3; a++) h[t][a] = fmaxf(fmaf(h[t][a], fmaf(alpha, pz[t][a] * pz[t][a], beta), c0[a]), 1e-10f);
            q[t][0] = fmaf(dcc_b, q[t][0], fmaf(dcc_a, pz[t][0] * pz[t][0], c1[0]));
            q[t][1] = fmaf(dcc_b, q[t][1], fmaf(dcc_a, pz[t][1] * pz[t][0], c1[1]));
            q[t][2] = fmaf(dcc_b, q[t][2], fmaf(dcc_a, pz[t][1] * pz[t][1], c1[2]));
            q[t][3] = fmaf(dcc_b, q[t][3], fmaf(dcc_a, pz[t][2] * pz[t][0], c1[3]));
            q[t][4] = fmaf(dcc_b, q[t][4], fmaf(dcc_a, pz[t][2] * pz[t][1], c1[4]));
            q[t][5] = fmaf(dcc_b, q[t][5], fmaf(dcc_a, pz[t][2] * pz[t][2], c1[5]));
          }
        }
        // chi-square mixing variable (calibration.py:565): nu = 4 / 5 straight-line, other integer dof through the general
        // product form, non-integer dof (calibrated joint models, calibration.py:399) by Marsaglia-Tsang - first attempt inline,
        // the ~3 % rejections out of line with their own Philox blocks, exactly as Trial::market_day draws them
        float u;
        if (kk == 4 || kk == 5) u = fmaf(nchi * kodd, nchi, M::neg2log_(M::u01(w[1].z) * M::u01(w[1].w)));
        else if (kk != 0) u = chi2_int_fast<R>(kk, w[1].z, w[1].w, nchi);
        else u = chi2_marsaglia_tsang<R, false>(nu, nchi, w[1].z, key, (uint32_t)g, STREAM_MKT1);
        const float scale = sqrt_nu * M::rsqrt_(fmaxf(u, 1e-12f));
        // risk-free rate of the day (engine.py:117-123, :516), regime in the sign bit
        const float rf = fminf(fmaxf(fmaf(0.5f, in, irx_base), 0.0f), 15.0f) * 0.01f;
        JF8 rec;
        rec.a3 = vn;
        rec.b3 = u2f_(f2u_(rf) | ((uint32_t)reg << 31));
#pragma unroll
        for (int t = 0; t < 2; t++) {
          float ih[3];
#pragma unroll
          for (int a = 0; a < 3; a++) ih[a] = M::rsqrt_(fmaxf(h[t][a], 1e-10f));
          const float sh0 = h[t][0] * ih[0], sh1 = h[t][1] * ih[1], sh2 = h[t][2] * ih[2];
          // factor of cov_t = diag(sqrt h) chol(R_t), R_t unit-diagonal (calibration.py:556-561)
          const float i0 = M::rsqrt_(fmaxf(q[t][0], 1e-12f)), i1 = M::rsqrt_(fmaxf(q[t][2], 1e-12f)), i2 = M::rsqrt_(fmaxf(q[t][5], 1e-12f));
          const float r10 = q[t][1] * i1 * i0, r20 = q[t][3] * i2 * i0, r21 = q[t][4] * i2 * i1;
          const float s11 = fmaxf(1.0f - r10 * r10, 0.0f);
          const float il11 = M::rsqrt_(fmaxf(s11, 1e-30f));
          const float l11 = s11 * il11;
          const float l21 = (r21 - r20 * r10) * il11;
          const float l22 = M::sqrt_(fmaxf(1.0f - r20 * r20 - l21 * l21, 0.0f));
          float zz[3];
          zz[0] = sh0 * z0;
          zz[1] = sh1 * (r10 * z0 + l11 * z1);
          zz[2] = sh2 * (r20 * z0 + l21 * z1 + l22 * z2);
          float x[3];
#pragma unroll
          for (int a = 0; a < 3; a++) {
            const float zs = t ? -zz[a] : zz[a];                  // antithetic twin: negated normal vector (calibration.py:563)
            x[a] = clipr(mu[a] + zs * scale, -0.95f, 4.0f);
            pz[t][a] = (x[a] - mu[a]) * ih[a];
            part[t][a] += x[a];                                    // eight-day partial sums, Kahan across them (below)
          }
          if (t == 0) { rec.a0 = x[0]; rec.a1 = x[1]; rec.a2 = x[2]; }
          else { rec.b0 = x[0]; rec.b1 = x[1]; rec.b2 = x[2]; }
        }
        jf_store(xs + g, rec);
        if ((g & 7) == 7 || g == G - 1) {
#pragma unroll
          for (int t = 0; t < 2; t++)
#pragma unroll
            for (int a = 0; a < 3; a++) {
              const float y = part[t][a] - cmp[t][a], s = sum[t][a] + y;
              cmp[t][a] = (s - sum[t][a]) - y; sum[t][a] = s; part[t][a] = 0.0f;
            }
        }
      }
#pragma unroll
      for (int t = 0; t < 2; t++)
#pragma unroll
        for (int a = 0; a < 3; a++) shift[t][a] = p.mm ? (p.mm_theo[a] - sum[t][a] / (R)G) * 0.1f : 0.0f;      // (moment matching off: the raw draws, calibration.py:579)
    }

    // ---- VIX burn-in: generated days before the window (engine.py:100-115)
    float vix[2] = {p.const_vix, p.const_vix};
    {
      int g = 0;
      for (; g + 4 <= so; g += 4) {
        JF8 v[4];
#pragma unroll
        for (int d = 0; d < 4; d++) v[d] = jf_load(xs + g + d);
#pragma unroll
        for (int d = 0; d < 4; d++) vix_day(p, fjs, g + d, v[d], shift, vix);
      }
      for (; g < so; g++) vix_day(p, fjs, g, jf_load(xs + g), shift, vix);
    }

    // ---- pass 2: the simulated window
    float liq = 0.0f, cred = 0.0f;
    float te_prev[2][NACT], bias[2][NACT], ddq[2][NACT], ddm[2][NACT], tq[2], tm[2], vmin[2], vmax[2];
    double cp[2][NACT], cp_tlt[2];
    int jump_days[2] = {0, 0}, reg_days1 = 0;
#pragma unroll
    for (int t = 0; t < 2; t++) {
      tq[t] = INFINITY; tm[t] = 1.0f; cp_tlt[t] = 1.0; vmin[t] = INFINITY; vmax[t] = -INFINITY;
#pragma unroll
      for (int a = 0; a < NACT; a++) { te_prev[t][a] = 0.0f; bias[t][a] = 0.0f; ddq[t][a] = INFINITY; ddm[t][a] = 1.0f; cp[t][a] = 1.0; }
    }
    uint32_t streams[1 + NACT];
    streams[0] = STREAM_STRESS;
#pragma unroll
    for (int a = 0; a < NACT; a++) streams[1 + a] = STREAM_TE0 + (uint32_t)p.act_id[a];
    const float ema = (float)(2.0 / 127.0);
    const bool want_tlt = p.bench_tlt != 0;
    const uint32_t amask = ALLSLOTS ? (uint32_t)p.asset_mask : ~0u;      // (compacted builds: every slot is live)

    const int64_t ecol = (pj - (ka.first_trial >> 1)) * 2;      // EMIT: this pair's first column (the chunk starts on an even trial)
    JF8 vnext;
    vnext.a0 = vnext.a1 = vnext.a2 = vnext.a3 = vnext.b0 = vnext.b1 = vnext.b2 = vnext.b3 = 0.0f;
    if (T > 0) vnext = jf_load(xs + so);
    for (int i = 0; i < T; i++) {
      const int g = so + i;
      const JF8 v = vnext;
      if (i + 1 < T) vnext = jf_load(xs + g + 1);
      const int reg = (int)(f2u_(v.b3) >> 31);
      const float rf = fabsf(v.b3);
      const DevFJ<R>& c = fjs[reg];
      reg_days1 += reg;

      // the day's Philox blocks, rounds interleaved: stress + one per tracking-error series
      U4 w[1 + NACT];
      rng_blocks<1 + NACT>(key, (uint32_t)g, streams, w);

      // latent stress channels (calibration.py:798-816): shared by the twins
      float ln, cn;
      M::box_muller(w[0].x, w[0].y, ln, cn);
      const float ju = (float)(w[0].z >> 8) * (float)(1.0 / 16777216.0);
      const float sign = (w[0].w & 1u) ? 1.0f : -1.0f;
      {
        const float l1 = clipx(c.liq_mu + c.liq_phi * (liq - c.liq_mu) + c.liq_sigma * ln, 0.0f, 3.0f);
        const float c1 = clipx(c.cr_mu + c.cr_phi * (cred - c.cr_mu) + c.cr_sigma * cn, 0.0f, 3.0f);
        liq = i == 0 ? c.liq_mu : l1;          // no draws on the first day (calibration.py:798-803)
        cred = i == 0 ? c.cr_mu : c1;
      }
      const float lq3 = clipx(liq, 0.0f, 3.0f), cr3 = clipx(cred, 0.0f, 3.0f);
      // financing spread, the part both twins share (engine.py:328-342, calibration.py:833-857)
      const float irx_level = rf * 100.0f;
      const float term = 1.25f - 1.1f * cr3;
      const float inv = fmaxf(-((irx_level + term) - irx_level), 0.0f);
      const float spd_sh = p.f_base + p.f_inv * inv + (p.f_liq * lq3 + p.f_cr * cr3);
      const float nz = fmaf(c.vnoise, vn_of(v), c.vdrift);
      const float jp_sh = fmaf(0.003f, liq, c.jbase);

      // tracking-error innovations (engine.py:185): shared by the twins
      float innov[NACT];
#pragma unroll
      for (int a = 0; a < NACT; a++) {
        if (ALLSLOTS && !((amask >> a) & 1)) { innov[a] = 0.0f; continue; }
        if (TEDF) innov[a] = student_t_block<R, false>(w[1 + a], p.act[a].te_df, p.act[a].te_fast_k, key, (uint32_t)g, streams[1 + a]);
        else innov[a] = ((K10 >> a) & 1) ? student_t10<R>(w[1 + a]) : student_t5<R>(w[1 + a]);
      }

      float x0[2], x1[2], x2[2];
      x0[0] = v.a0 + shift[0][0]; x1[0] = v.a1 + shift[0][1]; x2[0] = v.a2 + shift[0][2];
      x0[1] = v.b0 + shift[1][0]; x1[1] = v.b1 + shift[1][1]; x2[1] = v.b2 + shift[1][2];
      bool hit[2];
#pragma unroll
      for (int t = 0; t < 2; t++) {
        // VIX recursion (engine.py:100-115)
        const float shock = fmaxf(-x0[t], 0.0f) * c.vinvden;
        const float vj = c.vjscale * fmaxf(0.0f, shock - c.vjthr);
        const float cand = fmaf(c.vphi, vix[t], nz) + vj;
        vix[t] = g == 0 ? p.vix_base[reg] : fmaxf(cand, 10.0f);
        vmin[t] = fminf(vmin[t], vix[t]); vmax[t] = fmaxf(vmax[t], vix[t]);
        // jump probability (calibration.py:818-826)
        const float vamp = fmaxf((vix[t] - 25.0f) * (float)(1.0 / 30.0), 0.0f);
        const float jprob = clipx(fmaf(0.008f, vamp, jp_sh), 0.0f, 0.01f);
        hit[t] = i >= 1 && ju < jprob;
      }
      float e_ret[2][NACT];                      // EMIT: the day's layer returns of both twins
      float sjv[2] = {0.0f, 0.0f};
      if (hit[0] || hit[1]) {
        const float jump = jf_jump_t5(key.k0, key.k1, key.t0, key.t1, (uint32_t)g) * c.jscale;
#pragma unroll
        for (int t = 0; t < 2; t++)
          if (hit[t]) { sjv[t] = jump * sign; if (jump > 0.0f) jump_days[t]++; }
      }
#pragma unroll
      for (int t = 0; t < 2; t++) {
        const float spy = clipfast(x0[t] - sjv[t], -0.95f, 3.0f);
        const float qqq = clipfast(fmaf(-1.12f, sjv[t], x1[t]), -0.95f, 4.0f);
        const float tlt = x2[t];
        const float sv = fmaxf(vix[t] - 20.0f, 0.0f);
        const float spread = clipx(fmaf(p.f_vix, sv, spd_sh), p.f_min, p.f_max);
        const float carry = fmaxf(rf + spread, 0.0f);
        const float vmult = clipx(M::sqrt_(vix[t] * 0.05f), 0.5f, 1.5f);
        const float vrlive = i >= 1 ? vmult * c.rmult : 0.0f;      // no tracking error on day 0 (engine.py:184)
#pragma unroll
        for (int a = 0; a < NACT; a++) {
          if (ALLSLOTS && !((amask >> a) & 1)) { if (EMIT) e_ret[t][a] = 0.0f; continue; }
          const DevAsset<R>& ap = p.act[a];
          const float un = fmaf(ap.w_spy, spy, fmaf(ap.w_qqq, qqq, ap.w_tlt * tlt));
          const float lev = ap.lev * un - fmaf(ap.bl_252, carry, ap.er_daily);
          const float ds = un < 0.0f ? ap.te_down : 0.98f;
          const float mm = fmaf(0.3f, fabsf(un), 1.0f) * ds;
          const float te_raw = fmaf(ap.te_rho, te_prev[t][a], (innov[a] * vrlive) * (ap.te_scale * mm));
          te_prev[t][a] = te_raw;
          bias[t][a] = fmaf(ema, te_raw, (1.0f - ema) * bias[t][a]);
          float te = te_raw - bias[t][a];
          if (TEDF && ap.te_clip > 0.0f) te = ap.te_clip * M::tanh_(M::div_(te, ap.te_clip));
          const float er = clipfast(lev + te, -0.999f, 10.0f);
          cp[t][a] = ::fma(cp[t][a], (double)er, cp[t][a]);
          ddq[t][a] = fminf(1.0f, fmaf(ddq[t][a], er, ddq[t][a]));
          ddm[t][a] = fminf(ddm[t][a], ddq[t][a]);
          if (EMIT) e_ret[t][a] = er;
        }
        if (want_tlt) {
          cp_tlt[t] = ::fma(cp_tlt[t], (double)tlt, cp_tlt[t]);
          tq[t] = fminf(1.0f, fmaf(tq[t], tlt, tq[t]));
          tm[t] = fminf(tm[t], tq[t]);
        }
      }
      if (EMIT && ka.rec) {
        struct alignas(8) F2 { float a, b; };
        const int64_t ncol = ka.rec_ncol;
        F2* const base = (F2*)(ka.rec + (int64_t)i * ncol + ecol);
        const int64_t fstride = (int64_t)T * ncol / 2;          // fields are T * ncol floats apart (ncol is even)
#pragma unroll
        for (int a = 0; a < NACT; a++) { F2 v; v.a = e_ret[0][a]; v.b = e_ret[1][a]; base[(int64_t)a * fstride] = v; }
        { F2 v; v.a = u2f_(f2u_(vix[0]) | ((uint32_t)reg << 31)); v.b = u2f_(f2u_(vix[1]) | ((uint32_t)reg << 31)); base[(int64_t)NACT * fstride] = v; }
        { F2 v; v.a = (float)(cp[0][kSpySlot] * 100.0); v.b = (float)(cp[1][kSpySlot] * 100.0); base[(int64_t)(NACT + 1) * fstride] = v; }
      }
    }

    // ---- per-trial outputs (engine.py:578-603), validity as in Trial::run
    const int64_t N = ka.out.n_trials;
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int64_t col = pj * 2 + t - ka.first_trial;
      if (col < 0 || col >= ka.n_trials) continue;
      uint32_t flags = 0;
      bool bad = false;
      // a non-finite generated return poisons the column mean and with it every shifted return of the window
      if (!isfinite_(shift[t][0]) || !isfinite_(shift[t][1]) || !isfinite_(shift[t][2])) flags |= LETF_FLAG_NONFINITE_RET;
      if (T > 0 && (!(vmin[t] > 0.0f) || !isfinite_(vmax[t]))) flags |= LETF_FLAG_BAD_VIX;
#pragma unroll
      for (int a = 0; a < NACT; a++) if (T > 0 && ((amask >> a) & 1) && !(cp[t][a] > 0.0 && isfinite_(cp[t][a]))) bad = true;
      if (want_tlt && T > 0 && !(cp_tlt[t] > 0.0 && isfinite_(cp_tlt[t]))) bad = true;
      if (T == 0 || bad || !(vmin[t] >= 5.0f) || !(vmax[t] <= 120.0f)) flags |= LETF_FLAG_REPAIRED;
      for (int s = 0; s < NS; s++) {
        if (s >= p.n_strat) break;
        if (EMIT && p.strat[s].type != LETF_ST_BENCHMARK && p.strat[s].type != LETF_ST_BENCHMARK_TLT) continue;
        if (EMIT && p.strat[s].route < 0) continue;                       // routed rule-kernel builds (DevStrategy::route)
        const int64_t row = EMIT && p.strat[s].route > 0 ? p.strat[s].route - 1 : s;
        double c = 1.0; float qm = tm[t];
        if (p.strat[s].type == LETF_ST_BENCHMARK_TLT) c = cp_tlt[t];
        else {
          qm = 1.0f;
#pragma unroll
          for (int a = 0; a < NACT; a++) if (p.act_id[a] == p.strat[s].asset) { c = cp[t][a]; qm = ddm[t][a]; }
        }
        ka.out.final_wealth[row * N + col] = T > 0 ? (double)p.W0 * c : 0.0;
        ka.out.max_dd[row * N + col] = T > 1 ? 1.0 - (double)qm : 0.0;
        ka.out.num_trades[row * N + col] = 0;
        if (ka.out.journal_count) ka.out.journal_count[row * N + col] = 0;
        if (ka.out.journal_volume) ka.out.journal_volume[row * N + col] = 0.0;
      }
      if (ka.out.start_offset) ka.out.start_offset[col] = so;
      if (ka.out.flags) ka.out.flags[col] = flags;
      if (ka.out.regime_days) { ka.out.regime_days[col] = T - reg_days1; ka.out.regime_days[N + col] = reg_days1; }
      if (ka.out.jump_days) ka.out.jump_days[col] = jump_days[t];
    }
  }

  static LETF_HD float vn_of(const JF8& v) { return v.a3; }

  // one burn-in day: only the VIX recursion consumes it
  static LETF_HD void vix_day(const DevParams<R, NS, NA>& p, const DevFJ<R>* fjs, int g, const JF8& v, const float (&shift)[2][3], float (&vix)[2]) {
    const int reg = (int)(f2u_(v.b3) >> 31);
    const DevFJ<R>& c = fjs[reg];
    const float nz = fmaf(c.vnoise, v.a3, c.vdrift);
    const float xr[2] = {v.a0 + shift[0][0], v.b0 + shift[1][0]};
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const float shock = fmaxf(-xr[t], 0.0f) * c.vinvden;
      const float vj = c.vjscale * fmaxf(0.0f, shock - c.vjthr);
      const float cand = fmaf(c.vphi, vix[t], nz) + vj;
      vix[t] = g == 0 ? p.vix_base[reg] : fmaxf(cand, 10.0f);
    }
  }
};

// DevParams::fast_joint = (n_act << 8) | K10 mask; the patterns with a compiled-in pair kernel
LETF_HD constexpr int fast_joint_code(int nact, int k10) { return (nact << 8) | k10; }

}  // namespace letf
