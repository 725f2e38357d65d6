// letf_params.cuh — device-side parameter block (precision R) and its host-side construction
// from the public POD LetfParams.  Everything constant per run is folded here once.
#pragma once
#include "../../include/letf_b200.h"
#include <string.h>
#ifndef LETF_DEVICE_ONLY
#include <vector>
#endif
#include "letf_math.cuh"

namespace letf {

constexpr int kAssets = LETF_N_ASSETS;
constexpr int kPortMax = 2;      // portfolio-type strategies per launch
constexpr int kRingP = 200;      // SPY price history   (SMA 200 / 100, RSI 14)
constexpr int kRingS = 126;      // SPY return history  (momentum 126)
constexpr int kRingA = 60;       // signal-asset return history (vol 5/10/20/60, skew 60)

template <typename R> struct DevAsset {
  R lev, bl, er_daily, borrow_spread, te_base, te_df, te_rho, te_arw, te_down, te_clip, beta;
  R bl_252;          // max(leverage - 1, 0) / 252      (fp32 fast path of utils.py:169-216)
  R beta_eff;        // beta if it applies, else exactly 1
  R w_spy, w_qqq;    // beta_eff on the series' own underlying, 0 on the other (FAST loop: u = w_spy*spy + w_qqq*qqq)
  R w_tlt;           // same for TLT-based series (TMF); the pair kernel of letf_jfast.cuh adds w_tlt*tlt
  R carry_const;     // iid-normal market: expense/252 + borrow at the constant rate (loop invariant)
  R te_scale;        // te_base * sqrt(1 - rho^2)
  int und, beta_applies, te_fast_k, pad;
};

template <typename R> struct DevJoint {
  R mu[3], lv[3];
  R rb[6];            // lower triangle of nearest_psd(corr): 00 10 11 20 21 22
  R nu, alpha, beta, one_m_ab, dcc_a, dcc_b, one_m_dcc;
  // fp32 fast path: (1 - alpha - beta) * lv[a], (1 - dcc_a - dcc_b) * rbar[k], sqrt(nu) folded once per regime
  R c0[3], c1[6], sqrt_nu;
  int nu_fast_k, pad;
};

template <typename R> struct DevVix { R phi, drift /* (1-phi)*target */, noise_std, jump_thr, jump_scale, denom; };
// the per-regime constants the window loop of letf_jfast.cuh reads every day, packed for 128-bit constant loads
template <typename R> struct alignas(16) DevFJ {
  R vphi, vnoise, vdrift, vjthr;          // VIX recursion (engine.py:100-115)
  R vjscale, vinvden, rmult, jbase;       // 1 / max(daily_std, 1e-4); tracking-error regime multiplier (engine.py:179)
  R liq_mu, liq_phi, liq_sigma, jscale;   // latent stress (calibration.py:787-830)
  R cr_mu, cr_phi, cr_sigma, pad;
};
template <typename R> struct DevStress { R liq_mu, liq_phi, liq_sigma, cr_mu, cr_phi, cr_sigma, jbase, jscale; };

// legacy_hybrid block bootstrap (bootstrap.py:310-458, engine.py:50-80).  The candidate lists a block draw can
// select from - every block of a (pool, regime), or only those whose cumulative SPY return is >= 0 / < 0
// (momentum continuation, bootstrap.py:224-229) - are flattened on the host into (first history row, return) pairs.
constexpr int kMarketBoot = LETF_MARKET_BOOTSTRAP;
template <typename R> struct DevBoot {
  int block_size, block_min, block_max, hist_len;
  R log1mp;                     // ln(1 - 1 / block_mean): inverse CDF of the geometric block length
  R bias[2];
  R w1, nw1, w2, omw2, noise_w2;   // sampler weight, 1 - it; engine weight W, 1 - W, clip(1 - W, 0, 1)
  R ch1[2][9], ch2[2][9];       // Cholesky factors of the two noise blends, row-major
  int off[2][2][3], cnt[2][2][3];  // [pool A/B][regime][all / non-negative / negative]
  R sigma[2][3];                // pool B kernel bandwidth = std of the candidates' block returns (bootstrap.py:247-249)
  const R* hist[3];             // cleaned SPY, QQQ, TLT history
  const int* cstart;
  const R* cret;
};

template <typename R> struct DevStrategy {
  int type, asset, asset2, taxable, port;
  int route;      // output row of this slot: 0 = its own index, k > 0 = row k - 1, < 0 = slot switched off (routed typed builds, letf_api.cu)
  R w[kAssets];
  R p[LETF_STRAT_NPARAM];
};

template <typename R, int NS, int NA> struct DevParams {
  int market_mode, T, buffer, random_start, min_offset, offset_range, antithetic, mm, institutional,
      has_vix, has_stress, has_funding, asset_mask, n_strat, apply_costs, sig_asset,
      bench_mask /* assets tracked for a buy&hold strategy */, bench_tlt, need_tlt,
      fast_layers /* 0, or 1 + bitmask of the slots whose tracking error is t(10) (others t(5)): every slot active,
                     no tanh cap, no funding model, no TLT series -> straight-line day step (letf_fast.cuh) */,
      fast_joint /* 0, or (n_act << 8) | mask of the t(10) slots: institutional joint-t market with moment matching,
                    antithetic pairs, stress + funding models, every active series buy&hold -> one
                    antithetic pair per thread (letf_jfast.cuh); mask 0x80 = tracking-error dof / tanh cap read at run time */;
  R W0, sim_years;
  R start_p1;                   // P(start in regime 1), LetfParams::start_regime_prob1
  int flip[2];
  R log_pstay[2];
  int n_dur[2];
  const int32_t* dur[2];
  R reg_mean[2], reg_std[2];
  DevJoint<R> joint[2];
  R mm_theo[3];
  R pt_chol[2][9];
  R pt_df; int pt_fast_k, pad0;
  R n_mu[3], n_ch[6];      // normal market: lower-triangular Cholesky 00 10 11 20 21 22
  R const_vix, const_rf;
  DevVix<R> vix[2];
  R vix_base[2], irx_base[2];
  DevStress<R> stress[2];
  DevFJ<R> fj[2];
  R f_base, f_vix, f_inv, f_liq, f_cr, f_min, f_max;
  DevAsset<R> assets[kAssets];
  // the ETF series this launch computes: NA == 6 -> slot j is asset j (act_on = mask bit);
  // NA < 6 -> the active assets compacted into the first n_act slots (lean buy&hold builds)
  DevAsset<R> act[NA];
  int act_id[NA], act_on[NA], act_bench[NA];
  int n_act, pad_a;
  R cash_daily[2], spread_cost[2], rebalance_cost;
  DevStrategy<R> strat[NS];
  DevBoot<R> boot;
};


#ifndef LETF_DEVICE_ONLY
// ---- host: validate + convert -------------------------------------------------------------
// flat bootstrap tables in precision R (uploaded once per engine; the emulation build keeps them on the host)
template <typename R> struct BootTables {
  std::vector<R> hist;          // [3][hist_len]
  std::vector<int> cstart;
  std::vector<R> cret;
  DevBoot<R> meta;              // offsets / counts / bandwidths; pointers filled by the owner
};
inline const char* validate_params(const LetfParams& s) {
  if (s.abi_version != LETF_ABI_VERSION) return "abi_version mismatch";
  if (s.market_mode < 0 || (s.market_mode > 2 && s.market_mode != LETF_MARKET_BOOTSTRAP)) return "bad market_mode";
  if (s.market_mode == LETF_MARKET_BOOTSTRAP) {
    const LetfBootstrap& b = s.boot;
    if (b.hist_len <= 0 || !b.hist_spy || !b.hist_qqq || !b.hist_tlt) return "bootstrap history missing";
    if (!(b.block_min >= 1 && b.block_min <= b.block_max && b.block_max <= b.block_size)) return "bootstrap block lengths must satisfy 1 <= min <= max <= block_size";
    if (!(b.block_mean > 1.0)) return "bootstrap block_mean must be > 1";
    for (int q = 0; q < 2; q++) for (int r = 0; r < 2; r++) {
      const LetfBootstrapPool& pl = q ? b.pool_b[r] : b.pool_a[r];
      if (pl.n_blocks < 0 || (pl.n_blocks > 0 && (!pl.start || !pl.block_ret))) return "bootstrap pool table missing";
      for (int i = 0; i < pl.n_blocks; i++)
        if (pl.start[i] < 0 || pl.start[i] + b.block_size > b.hist_len) return "bootstrap block outside the history";
    }
  }
  if (s.sim_days < 0 || s.buffer_days < 0) return "negative day count";
  if (s.n_strategies < 0 || s.n_strategies > LETF_MAX_STRATEGIES) return "bad n_strategies";
  if (s.random_start && s.buffer_days <= 0) return "random_start needs buffer_days > 0";
  if (!(s.start_regime_prob1 >= 0.0 && s.start_regime_prob1 <= 1.0)) return "start_regime_prob1 must be in [0, 1]";
  if (s.market_mode == LETF_MARKET_JOINT_T)
    for (int r = 0; r < 2; r++) if (!(s.joint[r].nu >= 2.0)) return "joint nu must be >= 2";
  int sig_asset = -1, nport = 0;
  for (int k = 0; k < s.n_strategies; k++) {
    const LetfStrategyParams& q = s.strategies[k];
    const double* p = q.p;
    auto is = [&](int i, double v) { return p[i] == v; };
    switch (q.type) {
      case LETF_ST_BENCHMARK: case LETF_ST_BENCHMARK_TLT: break;
      case LETF_ST_SMA: case LETF_ST_SMA_BAND: if (!is(0, 200)) return "sma_period must be 200"; break;
      case LETF_ST_PORTFOLIO: ++nport; if (p[0] < 1) return "rebalance_freq < 1"; break;
      case LETF_ST_VOL_TARGETING: if (!is(1, 20)) return "vol_targeting lookback must be 20"; break;
      case LETF_ST_COMPOSITE: if (!is(0, 200) || !is(1, 14)) return "composite windows must be 200/14"; break;
      case LETF_ST_ADAPTIVE_VOL: if (!is(2, 20) || !is(3, 200)) return "adaptive_vol windows must be 20/200"; break;
      case LETF_ST_DOWNSIDE_VOL: if (!is(1, 20)) return "downside_vol lookback must be 20"; break;
      case LETF_ST_CONVEX_VOL: if (!is(2, 200)) return "convex_vol sma_period must be 200"; break;
      case LETF_ST_VOL_VELOCITY: case LETF_ST_VOL_MOM: case LETF_ST_SKEW_CONVEX: case LETF_ST_META_ENSEMBLE: break;
      case LETF_ST_REGIME_ASYMMETRIC: if (!is(4, 5) || !is(5, 60)) return "regime_asymmetric lookbacks must be 5/60"; break;
      case LETF_ST_SKEW_KELLY: if (!is(1, 60) || !is(2, 20)) return "skew_kelly lookbacks must be 60/20"; break;
      case LETF_ST_MOM_VOL_CONV: if (!is(1, 126) || !is(2, 10) || !is(3, 60)) return "mom_vol_convergence lookbacks must be 126/10/60"; break;
      case LETF_ST_CONVICTION: if (!is(1, 126) || !is(2, 20) || !is(3, 100)) return "conviction lookbacks must be 126/20/100"; break;
      default: return "unknown strategy type";
    }
    bool uses_sig = q.type >= LETF_ST_VOL_TARGETING && q.type <= LETF_ST_CONVICTION && q.type != LETF_ST_COMPOSITE;
    if (q.type != LETF_ST_PORTFOLIO && q.type != LETF_ST_BENCHMARK_TLT && (q.asset < 0 || q.asset >= kAssets)) return "bad strategy asset";
    if (q.type == LETF_ST_COMPOSITE && (q.asset2 < 0 || q.asset2 >= kAssets)) return "bad defensive asset";
    if (uses_sig) {
      if (sig_asset >= 0 && sig_asset != q.asset) return "all volatility-signal strategies must trade the same asset";
      sig_asset = q.asset;
    }
  }
  return nullptr;
}

inline uint32_t strategy_type_mask(const LetfParams& s) {
  uint32_t m = 0;
  for (int k = 0; k < s.n_strategies; k++) m |= 1u << s.strategies[k].type;
  return m;
}

template <typename R>
inline void build_boot_tables(const LetfParams& s, BootTables<R>& t) {
  const LetfBootstrap& b = s.boot;
  memset(&t.meta, 0, sizeof(t.meta));
  t.hist.clear(); t.cstart.clear(); t.cret.clear();
  if (s.market_mode != LETF_MARKET_BOOTSTRAP) return;
  DevBoot<R>& m = t.meta;
  m.block_size = b.block_size; m.block_min = b.block_min; m.block_max = b.block_max; m.hist_len = b.hist_len;
  m.log1mp = (R)::log(1.0 - 1.0 / b.block_mean);
  for (int r = 0; r < 2; r++) {
    m.bias[r] = (R)b.momentum_bias[r];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
      m.ch1[r][i * 3 + j] = (R)b.sampler_chol[r][i][j];
      m.ch2[r][i * 3 + j] = (R)b.engine_chol[r][i][j];
    }
  }
  m.w1 = (R)b.sampler_weight; m.nw1 = (R)(1.0 - b.sampler_weight);
  m.w2 = (R)b.engine_weight; m.omw2 = (R)(1.0 - b.engine_weight);
  { double nw = 1.0 - b.engine_weight; nw = nw < 0 ? 0 : (nw > 1 ? 1 : nw); m.noise_w2 = (R)nw; }
  t.hist.resize((size_t)3 * b.hist_len);
  for (int i = 0; i < b.hist_len; i++) {
    t.hist[i] = (R)b.hist_spy[i]; t.hist[(size_t)b.hist_len + i] = (R)b.hist_qqq[i];
    t.hist[(size_t)2 * b.hist_len + i] = (R)b.hist_tlt[i];
  }
  for (int q = 0; q < 2; q++) for (int r = 0; r < 2; r++) {
    const LetfBootstrapPool& pl = q ? b.pool_b[r] : b.pool_a[r];
    for (int kind = 0; kind < 3; kind++) {
      m.off[q][r][kind] = (int)t.cstart.size();
      double mean = 0; int nc = 0;
      for (int i = 0; i < pl.n_blocks; i++) {
        const double v = pl.block_ret[i];
        if (kind == 0 || (kind == 1 && v >= 0) || (kind == 2 && v < 0)) {
          t.cstart.push_back(pl.start[i]); t.cret.push_back((R)v); mean += v; nc++;
        }
      }
      m.cnt[q][r][kind] = nc;
      if (q == 1) {
        double var = 0, sg = 0.05;
        if (nc > 0) {
          mean /= nc;
          for (int i = 0; i < pl.n_blocks; i++) {
            const double v = pl.block_ret[i];
            if (kind == 0 || (kind == 1 && v >= 0) || (kind == 2 && v < 0)) var += (v - mean) * (v - mean);
          }
          sg = ::sqrt(var / nc);
          if (sg < 1e-8) sg = 0.05;
        }
        m.sigma[r][kind] = (R)sg;
      }
    }
  }
}

// shape code of the antithetic-pair kernel (letf_jfast.cuh), 0 = not admitted; computed from the public parameter block so
// that kernel selection (letf_api.cu) and the device parameter block agree
inline int fast_joint_shape(const LetfParams& s, int na) {
  if (na >= kAssets || s.market_mode != LETF_MARKET_JOINT_T || !s.antithetic_pairs) return 0;      // (moment matching off: same kernel, zero shift)
  if (!s.institutional || !s.has_stress || !s.has_funding || s.sim_days < 2) return 0;
  for (int k = 0; k < s.n_strategies; k++)
    if (s.strategies[k].type != LETF_ST_BENCHMARK && s.strategies[k].type != LETF_ST_BENCHMARK_TLT) return 0;
  int nact = 0, k10 = 0;
  bool tedf = false;        // some series has a non-default tracking-error dof or a tanh cap: run-time sampler build (code 0x80)
  for (int a = 0; a < kAssets; a++) {
    if (!((s.asset_mask >> a) & 1)) continue;
    const int fk = df_fast_k(s.assets[a].te_df);
    if ((fk != 5 && fk != 10) || s.assets[a].te_clip > 0.0) tedf = true;
    bool bench = false;
    for (int k = 0; k < s.n_strategies; k++) if (s.strategies[k].type == LETF_ST_BENCHMARK && s.strategies[k].asset == a) bench = true;
    if (!bench) return 0;
    if (fk == 10) k10 |= 1 << nact;
    nact++;
  }
  if (tedf) k10 = 0x80;
  if (nact < 1 || nact > na) return 0;
  return (nact << 8) | k10;
}

template <typename R, int NS, int NA>
inline void convert_params(const LetfParams& s, const int32_t* const dev_dur[2], DevParams<R, NS, NA>& d,
                           const DevBoot<R>* boot = nullptr) {
  memset(&d, 0, sizeof(d));
  if (boot) d.boot = *boot;
  d.market_mode = s.market_mode; d.T = s.sim_days; d.buffer = s.buffer_days; d.random_start = s.random_start;
  d.min_offset = s.buffer_days / 4 < 50 ? s.buffer_days / 4 : 50;
  d.offset_range = s.buffer_days - d.min_offset;
  d.antithetic = s.antithetic_pairs; d.mm = s.moment_matching && s.market_mode == LETF_MARKET_JOINT_T;
  d.institutional = s.institutional; d.has_vix = s.has_vix_dynamics; d.has_stress = s.has_stress;
  d.has_funding = s.has_funding && s.institutional; d.asset_mask = s.asset_mask; d.n_strat = s.n_strategies;
  d.apply_costs = s.apply_costs;
  d.W0 = (R)s.initial_capital; d.sim_years = (R)s.sim_years;
  d.start_p1 = (R)s.start_regime_prob1;
  d.sig_asset = 0;
  for (int r = 0; r < 2; r++) {
    d.flip[r] = s.transition[r][1 - r] > 0 ? 1 : 0;
    double ps = s.transition[r][r]; ps = ps < 0.80 ? 0.80 : (ps > 0.995 ? 0.995 : ps);
    d.log_pstay[r] = (R)::log(ps);
    d.n_dur[r] = s.n_duration[r]; d.dur[r] = dev_dur[r];
    d.reg_mean[r] = (R)s.regime_daily_mean[r]; d.reg_std[r] = (R)s.regime_daily_std[r];
    const LetfJointRegime& j = s.joint[r];
    DevJoint<R>& q = d.joint[r];
    double nu = j.nu, alpha = j.garch_alpha, beta = j.garch_beta;
    double tvs = nu > 2.0 ? (nu - 2.0) / nu : 0.5;
    for (int a = 0; a < 3; a++) {
      q.mu[a] = (R)j.mu[a];
      double lv = j.cov[a][a] * tvs; if (lv < 1e-8) lv = 1e-8;
      q.lv[a] = (R)lv;
    }
    double ea = nu > 2.0 ? alpha * nu / (nu - 2.0) : alpha * 2.0;
    if (ea + beta >= 1.0) { double b2 = 0.98 - ea; beta = 0.70 > b2 ? 0.70 : b2; }
    q.rb[0] = (R)j.rbar[0][0]; q.rb[1] = (R)j.rbar[1][0]; q.rb[2] = (R)j.rbar[1][1];
    q.rb[3] = (R)j.rbar[2][0]; q.rb[4] = (R)j.rbar[2][1]; q.rb[5] = (R)j.rbar[2][2];
    q.nu = (R)nu; q.alpha = (R)alpha; q.beta = (R)beta; q.one_m_ab = (R)(1 - alpha - beta);
    q.dcc_a = (R)j.dcc_a; q.dcc_b = (R)j.dcc_b; q.one_m_dcc = (R)(1 - j.dcc_a - j.dcc_b);
    q.nu_fast_k = df_fast_k(nu);
    for (int a = 0; a < 3; a++) q.c0[a] = (R)((1 - alpha - beta) * (double)q.lv[a]);
    for (int a = 0; a < 6; a++) q.c1[a] = (R)((1 - j.dcc_a - j.dcc_b) * (double)q.rb[a]);
    q.sqrt_nu = (R)::sqrt(nu);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) d.pt_chol[r][a * 3 + b] = (R)s.param_t_chol[r][a][b];
    const LetfVixRegime& v = s.vix[r];
    d.vix[r].phi = (R)v.phi; d.vix[r].drift = (R)((1 - v.phi) * v.target); d.vix[r].noise_std = (R)v.noise_std;
    d.vix[r].jump_thr = (R)v.jump_threshold; d.vix[r].jump_scale = (R)v.jump_scale;
    d.vix[r].denom = (R)(s.regime_daily_std[r] > 1e-4 ? s.regime_daily_std[r] : 1e-4);
    d.vix_base[r] = (R)s.vix_base[r]; d.irx_base[r] = (R)s.irx_base[r];
    const LetfStressRegime& t = s.stress[r];
    DevStress<R>& u = d.stress[r];
    u.liq_mu = (R)t.liq_mu; u.liq_phi = (R)t.liq_phi; u.liq_sigma = (R)t.liq_sigma;
    u.cr_mu = (R)t.credit_mu; u.cr_phi = (R)t.credit_phi; u.cr_sigma = (R)t.credit_sigma;
    u.jbase = (R)t.jump_base_prob; u.jscale = (R)t.jump_scale;
    d.cash_daily[r] = (R)(s.cash_rate[r] / 252); d.spread_cost[r] = (R)(s.base_spread_bps[r] / 10000);
    DevFJ<R>& f = d.fj[r];
    f.vphi = d.vix[r].phi; f.vnoise = d.vix[r].noise_std; f.vdrift = d.vix[r].drift; f.vjthr = d.vix[r].jump_thr;
    f.vjscale = d.vix[r].jump_scale; f.vinvden = (R)(1.0 / (s.regime_daily_std[r] > 1e-4 ? s.regime_daily_std[r] : 1e-4));
    f.rmult = r == 0 ? (R)1 : (R)1.15; f.jbase = u.jbase;
    f.liq_mu = u.liq_mu; f.liq_phi = u.liq_phi; f.liq_sigma = u.liq_sigma; f.jscale = u.jscale;
    f.cr_mu = u.cr_mu; f.cr_phi = u.cr_phi; f.cr_sigma = u.cr_sigma; f.pad = (R)0;
  }
  for (int a = 0; a < 3; a++) d.mm_theo[a] = (R)((s.joint[0].mu[a] + s.joint[1].mu[a]) / 2.0);
  d.pt_df = (R)s.param_t_df; d.pt_fast_k = df_fast_k(s.param_t_df);
  for (int a = 0; a < 3; a++) d.n_mu[a] = (R)s.normal_mu[a];
  d.n_ch[0] = (R)s.normal_chol[0][0]; d.n_ch[1] = (R)s.normal_chol[1][0]; d.n_ch[2] = (R)s.normal_chol[1][1];
  d.n_ch[3] = (R)s.normal_chol[2][0]; d.n_ch[4] = (R)s.normal_chol[2][1]; d.n_ch[5] = (R)s.normal_chol[2][2];
  d.const_vix = (R)s.const_vix; d.const_rf = (R)s.const_rf;
  d.f_base = (R)s.funding.base; d.f_vix = (R)s.funding.beta_vix; d.f_inv = (R)s.funding.beta_inv_curve;
  d.f_liq = (R)s.funding.beta_liquidity; d.f_cr = (R)s.funding.beta_credit;
  d.f_min = (R)s.funding.min_spread; d.f_max = (R)s.funding.max_spread;
  for (int a = 0; a < kAssets; a++) {
    const LetfAssetParams& x = s.assets[a];
    DevAsset<R>& y = d.assets[a];
    y.lev = (R)x.leverage; y.bl = (R)(x.leverage - 1.0); y.er_daily = (R)(x.expense_ratio / 252);
    y.borrow_spread = (R)x.borrow_spread; y.te_base = (R)x.te_base; y.te_df = (R)x.te_df; y.te_rho = (R)x.te_rho;
    double q = 1 - x.te_rho * x.te_rho; if (q < 1e-6) q = 1e-6;
    y.te_arw = (R)::sqrt(q); y.te_down = (R)x.te_downside; y.te_clip = (R)x.te_clip; y.beta = (R)x.beta;
    y.und = x.underlying; y.beta_applies = x.beta_applies; y.te_fast_k = df_fast_k(x.te_df);
    y.bl_252 = (R)(x.leverage > 1.0 ? (x.leverage - 1.0) / 252 : 0.0);
    y.beta_eff = x.beta_applies ? (R)x.beta : (R)1;
    y.w_spy = x.underlying == 0 ? y.beta_eff : (R)0;
    y.w_qqq = x.underlying == 1 ? y.beta_eff : (R)0;
    y.w_tlt = x.underlying == 2 ? y.beta_eff : (R)0;
    {
      double rate = s.const_rf + x.borrow_spread; if (!(rate > 0.0)) rate = 0.0;
      y.carry_const = (R)(x.expense_ratio / 252 + (x.leverage > 1.0 ? (x.leverage - 1.0) * rate / 252 : 0.0));
    }
    y.te_scale = (R)(x.te_base * ::sqrt(q));
  }
  d.rebalance_cost = (R)s.rebalance_cost;
  int nport = 0;
  for (int k = 0; k < s.n_strategies && k < NS; k++) {
    const LetfStrategyParams& x = s.strategies[k];
    DevStrategy<R>& y = d.strat[k];
    y.type = x.type; y.asset = x.asset; y.asset2 = x.asset2; y.taxable = x.taxable;
    y.port = x.type == LETF_ST_PORTFOLIO ? nport++ : -1;
    for (int a = 0; a < kAssets; a++) y.w[a] = (R)x.weights[a];
    for (int i = 0; i < LETF_STRAT_NPARAM; i++) y.p[i] = (R)x.p[i];
    bool uses_sig = x.type >= LETF_ST_VOL_TARGETING && x.type <= LETF_ST_CONVICTION && x.type != LETF_ST_COMPOSITE;
    if (uses_sig) d.sig_asset = x.asset;
    if (x.type == LETF_ST_BENCHMARK) d.bench_mask |= 1 << x.asset;
    if (x.type == LETF_ST_BENCHMARK_TLT) d.bench_tlt = 1;
  }
  d.need_tlt = d.bench_tlt;
  for (int a = 0; a < kAssets; a++) if (((s.asset_mask >> a) & 1) && s.assets[a].underlying == 2) d.need_tlt = 1;
  bool fast = (NA < kAssets) && !d.need_tlt && !d.has_funding && s.market_mode == LETF_MARKET_NORMAL &&
              !(s.random_start && s.buffer_days > 0);      // the straight-line loop keys its draws by the simulated day
  d.n_act = 0;
  for (int a = 0; a < kAssets; a++) {
    const bool on = (s.asset_mask >> a) & 1;
    if (on && ((d.assets[a].te_fast_k != 5 && d.assets[a].te_fast_k != 10) || d.assets[a].te_clip > (R)0 ||
               !((d.bench_mask >> a) & 1))) fast = false;
    if (NA == kAssets) {
      d.act[a] = d.assets[a]; d.act_id[a] = a; d.act_on[a] = on; d.act_bench[a] = (d.bench_mask >> a) & 1;
      d.n_act = kAssets;
    } else if (on && d.n_act < NA) {
      int j = d.n_act++;
      d.act[j] = d.assets[a]; d.act_id[j] = a; d.act_on[j] = 1; d.act_bench[j] = (d.bench_mask >> a) & 1;
    }
  }
  d.fast_layers = 0;
  if (fast && d.n_act >= 1) {
    int k10 = 0; bool no_te = true;
    for (int j = 0; j < d.n_act; j++) { if (d.act[j].te_fast_k == 10) k10 |= 1 << j; if (d.act[j].te_scale != (R)0) no_te = false; }
    d.fast_layers = (1 + k10) | (d.n_act << 8) | ((no_te ? 1 : 0) << 12);      // letf_fast.cuh fast_layers_code
  }
  d.fast_joint = fast_joint_shape(s, NA);
}
#endif  // LETF_DEVICE_ONLY

}  // namespace letf
