/*
 * letf_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, fp64, single-trial-at-a-time restatement of the reference's Monte
 * Carlo path (djbm10/LETF-Monte-Carlo, pure Python):
 *     letf/simulation/engine.py:423-619   simulate_single_path_fixed
 *     letf/calibration.py:480-590         simulate_joint_returns_t
 *     letf/calibration.py:787-857         stress state, borrow spread
 *     letf/strategy.py:55-1277            run_strategy_fixed (S1..S19)
 *     letf/trade.py:15-94                 TradeJournal
 * It consumes a LetfTape of pre-drawn variates so that
 * the live reference, run with a tape-replaying numpy Generator, and this file
 * see identical shocks.  Parity status: PINNED against outputs of the live
 * reference generated in the authoring container (oracle/make_golden.py ->
 * tests/golden/ *.json); the reference itself ships no golden vectors for this
 * path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (CUDA) path never
 * links or calls it.
 *
 * The pandas rolling kernels below restate pandas 3.0.2
 * (_libs/window/aggregations: Kahan add/remove sum+mean, Welford add/remove
 * variance with a from-scratch recompute when a removal collapses the variance, removals before
 * additions); tests/test_oracle_rolling.py checks them bit-for-bit against pandas, including
 * series with long runs of identical values (the clipped "downside" returns).
 */
#include "../include/letf_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define NA NAN

static double clipd(double x, double lo, double hi) {
  /* np.clip: NaN propagates */
  if (x != x) return x;
  return x < lo ? lo : (x > hi ? hi : x);
}
static double dmax(double a, double b) { return a > b ? a : b; } /* python max(a,b): b if b>a else a  -> use pymax */
static double pymax(double a, double b) { return (b > a) ? b : a; }

/* ------------------------------------------------------------------------ */
/* numpy pairwise summation (np.mean / np.sum on a contiguous float64 array) */
static double np_pairwise_sum(const double* a, long n) {
  if (n < 8) {
    double res = 0.0;
    for (long i = 0; i < n; i++) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    long i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    long n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
  }
}

/* ------------------------------------------------------------------------ */
/* pandas rolling kernels (fixed window w, min_periods mp, no NaNs in input) */

void orc_roll_sum(const double* x, int n, int w, int mp, double* out) {
  long nobs = 0, ncs = 0;
  double s = 0.0, ca = 0.0, cr = 0.0, prev = NA;
  for (int i = 0; i < n; i++) {
    if (i - w >= 0) {
      double v = x[i - w];
      nobs--;
      double y = -v - cr, t = s + y;
      cr = t - s - y;
      s = t;
    }
    double v = x[i];
    nobs++;
    double y = v - ca, t = s + y;
    ca = t - s - y;
    s = t;
    if (v == prev) ncs++; else ncs = 1;
    prev = v;
    if (nobs >= mp) out[i] = (ncs >= nobs) ? prev * (double)nobs : s;
    else out[i] = NA;
  }
}

void orc_roll_mean(const double* x, int n, int w, int mp, double* out) {
  long nobs = 0, ncs = 0, neg = 0;
  double s = 0.0, ca = 0.0, cr = 0.0, prev = NA;
  for (int i = 0; i < n; i++) {
    if (i - w >= 0) {
      double v = x[i - w];
      nobs--;
      double y = -v - cr, t = s + y;
      cr = t - s - y;
      s = t;
      if (v < 0) neg--;
    }
    double v = x[i];
    nobs++;
    double y = v - ca, t = s + y;
    ca = t - s - y;
    s = t;
    if (v < 0) neg++;
    if (v == prev) ncs++; else ncs = 1;
    prev = v;
    if (nobs >= mp && nobs > 0) {
      double r = s / (double)nobs;
      if (ncs >= nobs) r = prev;
      else if (neg == 0 && r < 0) r = 0.0;
      else if (neg == nobs && r > 0) r = 0.0;
      out[i] = r;
    } else out[i] = NA;
  }
}

/* rolling(w, min_periods=mp).std(), ddof=1.
 * pandas 3.0.2 roll_var as observed (the wheel ships no .pyx; behaviour established by black-box
 * probing, tests/test_oracle_rolling.py): Welford add/remove with Kahan-compensated mean updates,
 * removals before additions; and when a removal collapses the sum of squared deviations
 * (ssq_after < 1e3 * DBL_EPSILON * ssq_before, which includes going negative) the whole window is
 * recomputed from scratch.  No "all values identical" shortcut (older pandas had one). */
static void var_add(double v, long* nobs, double* mean, double* ssq, double* ca) {
  (*nobs)++;
  double pm = *mean - *ca;
  double y = v - *ca;
  double t = y - *mean;
  *ca = t + *mean - y;
  *mean = *mean + t / (double)(*nobs);
  *ssq = *ssq + (v - pm) * (v - *mean);
}
void orc_roll_std(const double* x, int n, int w, int mp, double* out) {
  long nobs = 0;
  double mean = 0.0, ssq = 0.0, ca = 0.0, cr = 0.0;
  for (int i = 0; i < n; i++) {
    if (i - w >= 0) {
      double v = x[i - w];
      double before = ssq;
      nobs--;
      if (nobs) {
        double pm = mean - cr;
        double y = v - cr;
        double t = y - mean;
        cr = t + mean - y;
        mean = mean - t / (double)nobs;
        ssq = ssq - (v - pm) * (v - mean);
      } else { mean = 0.0; ssq = 0.0; }
      if (ssq < 1e3 * 2.220446049250313e-16 * before) {
        nobs = 0; mean = 0.0; ssq = 0.0; ca = 0.0; cr = 0.0;
        for (int j = i - w + 1; j < i; j++) var_add(x[j], &nobs, &mean, &ssq, &ca);
      }
    }
    var_add(x[i], &nobs, &mean, &ssq, &ca);
    if (nobs >= mp && nobs > 1) {
      double r = ssq / (double)(nobs - 1);
      if (r < 0) r = 0.0;
      out[i] = sqrt(r);
    } else out[i] = NA;
  }
}

/* rolling(w, min_periods=mp).skew(): bias-corrected G1.  pandas 3.0 uses a
 * stable central-moment formulation; an online central-moment update agrees with
 * it to ~2e-14 absolute (checked), which is far below every threshold it feeds. */
void orc_roll_skew(const double* x, int n, int w, int mp, double* out) {
  long nobs = 0, ncs = 0;
  double mean = 0.0, m2 = 0.0, m3 = 0.0, prev = NA;
  for (int i = 0; i < n; i++) {
    if (i - w >= 0) {
      double v = x[i - w];
      long nn = nobs - 1;
      if (nn == 0) { mean = m2 = m3 = 0.0; }
      else {
        double delta = v - mean, dn = delta / (double)nn, term1 = delta * dn * (double)nobs;
        m3 = m3 - (term1 * dn * (double)(nobs + 1) - 3.0 * dn * m2);
        mean = mean - dn;
        m2 = m2 - term1;
      }
      nobs = nn;
    }
    double v = x[i];
    long nn = nobs + 1;
    double delta = v - mean, dn = delta / (double)nn, term1 = delta * dn * (double)nobs;
    m3 = m3 + term1 * dn * (double)(nn - 2) - 3.0 * dn * m2;
    m2 = m2 + term1;
    mean = mean + dn;
    nobs = nn;
    if (v == prev) ncs++; else ncs = 1;
    prev = v;
    if (nobs >= mp && nobs >= 3) {
      double d = (double)nobs;
      double B = m2 / d;
      if (ncs >= nobs) out[i] = 0.0;
      else if (B <= 1e-14) out[i] = NA;
      else out[i] = sqrt(d * (d - 1.0)) / (d - 2.0) * (m3 / d) / (B * sqrt(B));
    } else out[i] = NA;
  }
}

static void shift1(const double* in, int n, double* out) {
  for (int i = n - 1; i >= 1; i--) out[i] = in[i - 1];
  if (n > 0) out[0] = NA;
}

/* ------------------------------------------------------------------------ */
/* TradeJournal (trade.py:15-73) */
typedef struct {
  double pos[LETF_N_ASSETS];
  long count;
  double volume;
  /* optional full log */
  int cap, n;
  int* day; int* asset; int* action; double* shares; double* price;
} Journal;

static void journal_log(Journal* j, int day, int asset, double new_alloc, double portfolio_value,
                        double price) {
  if (price <= 0) return;
  double target_shares = new_alloc * portfolio_value / price;
  double change = target_shares - j->pos[asset];
  if (fabs(change) < 0.001) return;
  double shares = change > 0 ? change : fabs(change);
  j->count++;
  j->volume += shares * price;
  if (j->day && j->n < j->cap) {
    j->day[j->n] = day; j->asset[j->n] = asset; j->action[j->n] = change > 0 ? 1 : -1;
    j->shares[j->n] = shares; j->price[j->n] = price; j->n++;
  }
  j->pos[asset] = target_shares;
}

/* ------------------------------------------------------------------------ */
/* per-trial simulated frame (engine.py:537-550) */
typedef struct {
  int T;
  double* ret[LETF_N_ASSETS];
  double* price[LETF_N_ASSETS];
  double* cash;
  double* tlt_ret;
  double* tlt_price;
  double* vix;
  int* regime;
} Frame;

static void max_drawdown_and_final(const double* ec, int T, double* fw, double* mdd) {
  /* engine.py:578-586 */
  *fw = T > 0 ? ec[T - 1] : 0.0;
  double m = 0.0;
  if (T > 1) {
    double rm = ec[0], mn = INFINITY;
    for (int i = 0; i < T; i++) {
      if (ec[i] > rm || rm != rm) rm = ec[i]; /* np.maximum.accumulate (propagates NaN; not expected) */
      double dd = (ec[i] - rm) / (rm > 0 ? rm : 1.0);
      if (dd < mn || dd != dd) mn = dd;
    }
    m = fabs(mn);
  }
  *mdd = m;
}

/* strategy.py:8-53 with trade_size_pct = 0 */
static double txn_cost(const LetfParams* p, double daily_ret, int regime, double leverage) {
  double spread_cost = p->base_spread_bps[regime] / 10000;
  double rebalance_cost = p->rebalance_cost * leverage * fabs(daily_ret);
  return spread_cost + rebalance_cost + 0;
}

/* one strategy on one frame: returns equity curve in ec[T]; num_trades */
static int run_strategy(const LetfParams* p, const LetfStrategyParams* sp, const Frame* f,
                        double* ec, Journal* jr, double* wk /* >= 12*T doubles scratch */) {
  const int T = f->T;
  const double W0 = p->initial_capital;
  const double V = sqrt(252.0);
  int num_trades = 0;
  const double* r = (sp->asset >= 0 && sp->asset < LETF_N_ASSETS) ? f->ret[sp->asset] : NULL;
  const double* aprice = (sp->asset >= 0 && sp->asset < LETF_N_ASSETS) ? f->price[sp->asset] : NULL;
  const double* P = f->price[4];   /* SPY_Price */
  const double* spy_ret = f->ret[4];
  const double* cash = f->cash;
  double* a0 = wk, *a1 = wk + T, *a2 = wk + 2 * T, *a3 = wk + 3 * T, *a4 = wk + 4 * T,
        *a5 = wk + 5 * T, *a6 = wk + 6 * T;
  const int apply_costs = p->apply_costs;

#define VOLSTEP(ALLOC, THRESH, STICKY, CASHLEG)                                         \
  do {                                                                                   \
    double alloc_ = (ALLOC);                                                             \
    double turnover = fabs(alloc_ - prev_alloc), cost = 0.0;                             \
    if (turnover > (THRESH)) {                                                           \
      num_trades++;                                                                      \
      if (jr) journal_log(jr, i, sp->asset, alloc_, ec[i - 1], aprice[i]);               \
      cost = turnover * 0.0003;                                                          \
      prev_alloc = alloc_;                                                               \
    } else {                                                                             \
      if (STICKY) alloc_ = prev_alloc; else prev_alloc = alloc_;                         \
    }                                                                                    \
    double ret_ = (CASHLEG) ? (alloc_ * r[i]) + ((1 - alloc_) * cash[i]) : r[i] * alloc_;\
    if (apply_costs) ret_ -= cost;                                                       \
    ec[i] = ec[i - 1] * (1 + ret_);                                                      \
  } while (0)

  switch (sp->type) {
    case LETF_ST_BENCHMARK:
    case LETF_ST_BENCHMARK_TLT: {
      const double* rr = sp->type == LETF_ST_BENCHMARK ? r : f->tlt_ret;
      double cp = 1.0;
      for (int i = 0; i < T; i++) { cp = (i == 0) ? (1 + rr[0]) : cp * (1 + rr[i]); ec[i] = W0 * cp; }
      return 0;
    }
    case LETF_ST_SMA:
    case LETF_ST_SMA_BAND: {
      int period = (int)sp->p[0];
      double band = sp->type == LETF_ST_SMA_BAND ? sp->p[1] : 0.0;
      orc_roll_mean(P, T, period, 1, a0);
      shift1(a0, T, a1); /* sma_prev */
      int* pos = (int*)calloc(T > 0 ? T : 1, sizeof(int));
      for (int i = 1; i < T; i++) {
        double pp = P[i - 1], thr = sp->type == LETF_ST_SMA ? a1[i] : a1[i] * (1 - band);
        int buy = pp >= thr, sell = pp < thr; /* NaN -> both False */
        if (pos[i - 1] == 0) pos[i] = buy ? 1 : 0; else pos[i] = sell ? 0 : 1;
      }
      long nt = 0;
      for (int i = 1; i < T; i++) nt += abs(pos[i] - pos[i - 1]);
      num_trades = (int)nt;
      double lev = p->assets[sp->asset].leverage;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double ret = pos[i] == 1 ? r[i] : cash[i];
        if (apply_costs && abs(pos[i] - pos[i - 1]) > 0) ret -= txn_cost(p, r[i], f->regime[i], lev);
        ec[i] = ec[i - 1] * (1 + ret);
      }
      free(pos);
      return num_trades;
    }
    case LETF_ST_PORTFOLIO: {
      int freq = (int)sp->p[0];
      double pos[LETF_N_ASSETS], emb[LETF_N_ASSETS];
      int held[LETF_N_ASSETS], nh = 0;
      for (int a = 0; a < LETF_N_ASSETS; a++) if (sp->weights[a] != 0.0) held[nh++] = a;
      /* dict order of config['assets'] == ascending asset index for the reference's S6 (TQQQ, TMF) */
      for (int h = 0; h < nh; h++) { int a = held[h]; pos[a] = W0 * sp->weights[a]; emb[a] = p->assets[a].leverage; }
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        for (int h = 0; h < nh; h++) {
          int a = held[h];
          double ret = f->ret[a][i];
          pos[a] = pos[a] * (1 + ret);
          double tl = p->assets[a].leverage;
          if (tl > 1.0) {
            double u = ret / tl;
            if (fabs(1 + tl * u) > 0.01) emb[a] = tl * (1 + u) / (1 + tl * u); else emb[a] = tl;
          } else emb[a] = 1.0;
        }
        double total = 0.0; /* python sum(): 0 + v0 + v1 ... */
        for (int h = 0; h < nh; h++) total += pos[held[h]];
        ec[i] = total;
        if (i % freq == 0) {
          double cw[LETF_N_ASSETS];
          for (int h = 0; h < nh; h++) cw[held[h]] = pos[held[h]] / total;
          double wt = 0.0;
          for (int h = 0; h < nh; h++) wt += fabs(cw[held[h]] - sp->weights[held[h]]);
          double lt = 0.0;
          for (int h = 0; h < nh; h++) {
            int a = held[h]; double tl = p->assets[a].leverage;
            lt += fabs(emb[a] - tl) / tl * cw[a];
          }
          double turn = wt + lt;
          if (apply_costs && turn > 0.01) {
            double rc = turn * p->rebalance_cost * total;
            total -= rc;
            ec[i] = total;
          }
          for (int h = 0; h < nh; h++) { int a = held[h]; pos[a] = total * sp->weights[a]; emb[a] = p->assets[a].leverage; }
          num_trades += nh;
        }
      }
      return num_trades;
    }
    case LETF_ST_VOL_TARGETING: {
      double target = sp->p[0]; int lb = (int)sp->p[1];
      orc_roll_std(r, T, lb, lb, a0);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double cv = a0[i] * V, ps;
        if (cv != cv || cv < 0.01) ps = 1.0; else ps = clipd(target / cv, 0.2, 2.0);
        VOLSTEP(ps, 0.0001, 0, 0);
      }
      return num_trades;
    }
    case LETF_ST_COMPOSITE: {
      int sma_p = (int)sp->p[0], rsi_p = (int)sp->p[1]; double vix_th = sp->p[2];
      orc_roll_mean(P, T, sma_p, sma_p, a0); /* sma (unshifted, read at i-1) */
      for (int i = 0; i < T; i++) {
        double d = i == 0 ? NA : P[i] - P[i - 1];
        a1[i] = (d > 0) ? d : 0.0;            /* delta.where(delta > 0, 0) */
        a2[i] = -((d < 0) ? d : 0.0);         /* -delta.where(delta < 0, 0) (may be -0.0) */
      }
      orc_roll_mean(a1, T, rsi_p, rsi_p, a3);
      orc_roll_mean(a2, T, rsi_p, rsi_p, a4);
      int cur = 0; /* 0 CASH, 1 SAFE, 2 RISKY */
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double rs = a3[i - 1] / a4[i - 1];
        double rsi = 100 - (100 / (1 + rs));
        int score = 0;
        if (P[i - 1] > a0[i - 1]) score++;
        if (40 < rsi && rsi < 80) score++;
        if (f->vix[i - 1] < vix_th) score++;
        double ret; int target;
        if (score == 3) { ret = f->ret[sp->asset][i]; target = 2; }
        else if (score == 2) { ret = f->ret[sp->asset2][i]; target = 1; }
        else { ret = cash[i]; target = 0; }
        if (target != cur) {
          num_trades++;
          if (jr) {
            int ta = target == 2 ? sp->asset : (target == 1 ? sp->asset2 : 4 /* 'SPY' */);
            double na = target == 0 ? 0.0 : 1.0;
            journal_log(jr, i, ta, na, ec[i - 1], f->price[ta][i]);
          }
          cur = target;
        }
        ec[i] = ec[i - 1] * (1 + ret);
      }
      return num_trades;
    }
    case LETF_ST_ADAPTIVE_VOL: {
      double bull = sp->p[0], bear = sp->p[1]; int lb = (int)sp->p[2], sp_ = (int)sp->p[3];
      orc_roll_std(r, T, lb, lb, a0); shift1(a0, T, a1);
      orc_roll_mean(P, T, sp_, sp_, a2); shift1(a2, T, a3);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double cv = a1[i] * V, cs = a3[i];
        if (cv != cv || cs != cs || cv < 0.001) { ec[i] = ec[i - 1]; continue; }
        double tv = (P[i - 1] > cs) ? bull : bear;
        VOLSTEP(clipd(tv / cv, 0.0, 1.0), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_DOWNSIDE_VOL: {
      double target = sp->p[0]; int lb = (int)sp->p[1];
      for (int i = 0; i < T; i++) a0[i] = r[i] < 0 ? r[i] : 0.0;
      orc_roll_std(a0, T, lb, lb, a1); shift1(a1, T, a2);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double dv = a2[i] * V, al;
        if (dv != dv || dv < 0.001) al = 1.0; else al = clipd(target / dv, 0.0, 1.5);
        VOLSTEP(al, 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_CONVEX_VOL: {
      double target = sp->p[0], pw = sp->p[1]; int sma_p = (int)sp->p[2];
      orc_roll_std(r, T, 20, 1, a0); shift1(a0, T, a1);
      orc_roll_mean(P, T, sma_p, 1, a2); shift1(a2, T, a3);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double v = a1[i] * V;
        if (v != v || v < 0.001) v = 0.5;
        double al = target / v;
        if (P[i - 1] > a3[i]) al = pow(al, pw);
        VOLSTEP(clipd(al, 0.0, 1.0), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_VOL_VELOCITY: {
      double target = sp->p[0];
      orc_roll_std(r, T, 20, 1, a0); shift1(a0, T, a1);
      orc_roll_std(r, T, 5, 1, a2); shift1(a2, T, a3);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double ev = pymax(a1[i] * V, a3[i] * V);
        if (ev != ev || ev < 0.001) ev = 0.5;
        VOLSTEP(clipd(target / ev, 0.0, 1.0), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_VOL_MOM: {
      double target = sp->p[0];
      orc_roll_std(r, T, 20, 1, a0); shift1(a0, T, a1);
      for (int i = 0; i < T; i++) a1[i] = a1[i] * V;
      for (int i = 0; i < T; i++) a2[i] = i >= 5 ? a1[i] / a1[i - 5] - 1 : NA; /* pct_change(5), no fill */
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double v = a1[i], vm = a2[i];
        if (v != v || v < 0.001) v = 0.5;
        double al = target / v;
        if (vm == vm && vm < -0.10) al *= 1.2;
        if (vm == vm && vm > 0.10) al *= 0.7;
        VOLSTEP(clipd(al, 0.0, 1.0), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_SKEW_CONVEX: {
      double target = sp->p[0];
      orc_roll_skew(r, T, 60, 1, a0); shift1(a0, T, a1);
      orc_roll_std(r, T, 20, 1, a2); shift1(a2, T, a3);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double v = a3[i] * V, s = a1[i];
        if (v != v || v < 0.001) v = 0.5;
        double al = target / v;
        if (s == s && s > 0) al = pow(al, 1.3);
        else if (s == s && s < -0.5) al *= 0.5;
        VOLSTEP(clipd(al, 0.0, 1.0), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_META_ENSEMBLE: {
      double target = sp->p[0];
      for (int i = 0; i < T; i++) a0[i] = r[i] < 0 ? r[i] : 0.0;
      orc_roll_std(a0, T, 20, 1, a1); shift1(a1, T, a2);           /* d_vol */
      orc_roll_mean(P, T, 200, 1, a1); shift1(a1, T, a3);          /* sma */
      orc_roll_std(r, T, 5, 1, a1); shift1(a1, T, a4);             /* v_fast */
      orc_roll_std(r, T, 20, 1, a1); shift1(a1, T, a5);            /* v_slow */
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double dv = a2[i] * V;
        if (dv != dv || dv < 0.001) dv = 0.25;
        double al = target / dv;
        if (P[i - 1] > a3[i]) al = pow(al, 1.2);
        if (a4[i] * V > 1.5 * (a5[i] * V)) al *= 0.5;
        VOLSTEP(clipd(al, 0.0, 1.0), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_REGIME_ASYMMETRIC: {
      double base = sp->p[0], crisis = sp->p[1], alarm = sp->p[2], vthr = sp->p[3];
      int lf = (int)sp->p[4], ls = (int)sp->p[5];
      orc_roll_std(r, T, lf, 1, a0); shift1(a0, T, a1);
      orc_roll_std(r, T, ls, 1, a0); shift1(a0, T, a2);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double vf = a1[i] * V, vs = a2[i] * V, vr = vf / vs;
        int cm = (f->vix[i] > alarm) || (vr > vthr);
        double tv = cm ? crisis : base, al;
        if (vf != vf || vf < 0.001) al = 0.5; else al = clipd(tv / vf, 0.0, 1.2);
        VOLSTEP(al, 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_SKEW_KELLY: {
      double base = sp->p[0]; int skl = (int)sp->p[1], vl = (int)sp->p[2]; double kf = sp->p[3];
      orc_roll_std(r, T, vl, 1, a0); shift1(a0, T, a1);            /* realized vol */
      orc_roll_skew(r, T, skl, 1, a0); shift1(a0, T, a2);          /* skew */
      for (int i = 0; i < T; i++) a0[i] = r[i] < 0 ? r[i] : 0.0;
      orc_roll_std(a0, T, vl, 1, a3); shift1(a3, T, a4);           /* downside vol */
      orc_roll_mean(r, T, skl, 1, a3); shift1(a3, T, a5);          /* rolling mean */
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double vol = a1[i] * V, dv = a4[i] * V, sk = a2[i], mr = a5[i] * 252;
        if (vol != vol || vol < 0.001) vol = 0.25;
        if (dv != dv || dv < 0.001) dv = vol * 0.6;
        if (sk != sk) sk = 0.0;
        if (mr != mr) mr = 0.08;
        double ev, pen;
        if (sk < -0.5) { ev = dv * 1.5; pen = 0.6; }
        else if (sk < 0) { ev = dv * 1.2; pen = 0.8; }
        else if (sk > 0.5) { ev = vol * 0.9; pen = 1.2; }
        else { ev = vol; pen = 1.0; }
        double ks;
        if (ev > 0.01) { ks = (mr - 0.03) / (ev * ev); ks = ks * kf; ks = clipd(ks, 0.2, 2.0); }
        else ks = 1.0;
        double raw = (base / ev) * pen * (ks / 1.5);
        VOLSTEP(clipd(raw, 0.0, 1.5), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_MOM_VOL_CONV: {
      double base = sp->p[0]; int ml = (int)sp->p[1], vfl = (int)sp->p[2], vsl = (int)sp->p[3];
      double mthr = sp->p[4];
      orc_roll_sum(spy_ret, T, ml, 1, a0); shift1(a0, T, a1);
      orc_roll_std(r, T, vfl, 1, a0); shift1(a0, T, a2);
      orc_roll_std(r, T, vsl, 1, a0); shift1(a0, T, a3);
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double mom = a1[i], vf = a2[i] * V, vs = a3[i] * V;
        if (mom != mom) mom = 0.0;
        if (vf != vf || vf < 0.001) vf = 0.30;
        if (vs != vs || vs < 0.001) vs = 0.25;
        double mm = mom > mthr ? 1.3 : (mom > 0 ? 1.0 : 0.5);
        double ratio = vf / vs, vm, ev;
        if (ratio < 0.8) { vm = 1.2; ev = vf; }
        else if (ratio > 1.3) { vm = 0.6; ev = vf; }
        else { vm = 1.0; ev = vs; }
        double comb = mm * vm;
        double adj = base * comb;
        VOLSTEP(clipd(adj / ev, 0.0, 1.5), 0.0001, 0, 1);
      }
      return num_trades;
    }
    case LETF_ST_CONVICTION: {
      double base = sp->p[0]; int ml = (int)sp->p[1], vl = (int)sp->p[2], ts = (int)sp->p[3];
      double thr = sp->p[4];
      orc_roll_sum(spy_ret, T, ml, 1, a0); shift1(a0, T, a1);      /* momentum */
      for (int i = 0; i < T; i++) a0[i] = r[i] < 0 ? r[i] : 0.0;
      orc_roll_std(a0, T, vl, 1, a2); shift1(a2, T, a3);           /* downside vol */
      orc_roll_std(r, T, 5, 1, a2); shift1(a2, T, a4);             /* vol fast */
      orc_roll_std(r, T, 60, 1, a2); shift1(a2, T, a5);            /* vol slow */
      orc_roll_mean(P, T, ts, 1, a2); shift1(a2, T, a6);           /* trend sma */
      double prev_alloc = 0.0;
      if (T > 0) ec[0] = W0;
      for (int i = 1; i < T; i++) {
        double mom = a1[i], dv = a3[i] * V, vf = a4[i] * V, vs = a5[i] * V;
        if (mom != mom) mom = 0.0;
        if (dv != dv || dv < 0.001) dv = 0.20;
        if (vf != vf || vf < 0.001) vf = 0.30;
        if (vs != vs || vs < 0.001) vs = 0.25;
        double ms = mom > 0.15 ? 1.0 : (mom > 0.05 ? 0.7 : (mom > 0 ? 0.4 : 0.0));
        double tsc = P[i - 1] > a6[i] ? 0.5 : 0.0;
        double ratio = vf / vs;
        double vsc = ratio < 0.9 ? 0.5 : (ratio < 1.2 ? 0.3 : 0.0);
        double conv = ms + tsc + vsc;
        double ba = base / dv;
        double cmul = 0.3 + (conv * 0.55);
        double al = clipd(ba * cmul, 0.0, 1.5);
        VOLSTEP(al, thr, 1, 1);
      }
      return num_trades;
    }
    default: {
      /* strategy.py:1273-1277 default: SPY buy & hold */
      double cp = 1.0;
      for (int i = 0; i < T; i++) { cp = (i == 0) ? (1 + spy_ret[0]) : cp * (1 + spy_ret[i]); ec[i] = W0 * cp; }
      return 0;
    }
  }
#undef VOLSTEP
}

/* ------------------------------------------------------------------------ */
typedef struct LetfOracleSeries {
  /* optional per-day dump of ONE trial (debugging / layer-level parity) */
  int32_t* regime_full;  /* [G] */
  double* gen[5];        /* SPY, QQQ, TLT, VIX, IRX over G generated days */
  double* asset_ret[LETF_N_ASSETS]; /* [T] */
  double* liq; double* cred; double* jump; /* [T] */
  double* equity;        /* [S][T] */
} LetfOracleSeries;

static void chol3(const double c[3][3], double L[3][3]) {
  memset(L, 0, 9 * sizeof(double));
  L[0][0] = sqrt(c[0][0]);
  L[1][0] = c[1][0] / L[0][0];
  L[2][0] = c[2][0] / L[0][0];
  L[1][1] = sqrt(c[1][1] - L[1][0] * L[1][0]);
  L[2][1] = (c[2][1] - L[2][0] * L[1][0]) / L[1][1];
  L[2][2] = sqrt(c[2][2] - L[2][0] * L[2][0] - L[2][1] * L[2][1]);
}

/* ------------------------------------------------------------------------ */
/* legacy_hybrid block bootstrap: BlockBootstrapReturns.sample_returns (bootstrap.py:310-458)
 * followed by the second noise blend of generate_fat_tailed_returns (engine.py:50-80).
 * Tape ledger per block k: boot_len[k] = rng.geometric(1 / BLOCK_MEAN) (bootstrap.py:304);
 * boot_u[k][0..5] = momentum coin A (:226), block index A (:262), sub-section start A (:352),
 * momentum coin B, block index B (uniform :262 or weighted rng.choice :259), sub-section start B (:370).
 * rng.integers(0, m) on a tape uniform u is floor(u * m); rng.choice(m, p=w) is numpy's
 * cdf.searchsorted(u, side='right').  Empty pools (the reference's synthetic-block fallback,
 * bootstrap.py:214-222) are rejected by the parameter builder and not restated here. */
static int boot_pick(const LetfBootstrapPool* pool, int have_last, double last_ret, double bias,
                     double u_mom, double u_idx, int use_target, double target) {
  /* returns the index (into the pool's own block list) of the chosen block */
  const int nb = pool->n_blocks;
  int* cand = (int*)malloc(sizeof(int) * (size_t)(nb > 0 ? nb : 1));
  int nc = 0;
  int all = 1;
  if (have_last) {
    const int desired = last_ret >= 0 ? 1 : -1;
    int ns = 0;
    for (int i = 0; i < nb; i++) {
      const double r = pool->block_ret[i];
      if ((r >= 0 && desired >= 0) || (r < 0 && desired < 0)) cand[ns++] = i;
    }
    if (ns > 0 && bias > 0 && u_mom < bias) { nc = ns; all = 0; }
  }
  if (all) { for (int i = 0; i < nb; i++) cand[i] = i; nc = nb; }
  int pick;
  if (use_target && nc > 5) {
    /* Gaussian kernel weights around the economy block's SPY return (bootstrap.py:232-259) */
    double mean = 0, var = 0;
    for (int i = 0; i < nc; i++) mean += pool->block_ret[cand[i]];
    mean /= nc;
    for (int i = 0; i < nc; i++) { double d = pool->block_ret[cand[i]] - mean; var += d * d; }
    double sigma = sqrt(var / nc);
    if (sigma < 1e-8) sigma = 0.05;
    double W = 0;
    for (int i = 0; i < nc; i++) { double d = (pool->block_ret[cand[i]] - target) / sigma; W += exp(-0.5 * (d * d)); }
    double cum = 0; int idx = 0;
    for (int i = 0; i < nc; i++) {
      double d = (pool->block_ret[cand[i]] - target) / sigma;
      cum += exp(-0.5 * (d * d)) / W;
      if (cum <= u_idx) idx = i + 1;
    }
    if (idx > nc - 1) idx = nc - 1;
    pick = cand[idx];
  } else {
    int idx = (int)floor(u_idx * nc);
    if (idx > nc - 1) idx = nc - 1;
    pick = cand[idx];
  }
  free(cand);
  return pick;
}

static int boot_generate(const LetfParams* p, const LetfTape* tp, int64_t ti, int64_t n, const int* regf, int G,
                         double* spy, double* qqq, double* tlt) {
  const LetfBootstrap* b = &p->boot;
  if (!tp->boot_len || !tp->boot_u || !tp->boot_t) return LETF_ERR_INVALID;
  int cur = 0, k = 0, have_last = 0;
  double last_a = 0, last_b = 0;
  while (cur < G) {
    if (k >= tp->max_blocks) return LETF_ERR_INVALID;
    const int remaining = G - cur;
    int len = tp->boot_len[(int64_t)k * n + ti];
    if (len < b->block_min) len = b->block_min;
    if (len > b->block_max) len = b->block_max;
    if (len > remaining) len = remaining;
    int ones = 0;
    for (int j = 0; j < len; j++) ones += regf[cur + j];
    const int regime = 2 * ones > len ? 1 : 0;          /* int(np.median(block_regime_path)) */
    const double bias = b->momentum_bias[regime];
#define BU(s) tp->boot_u[((int64_t)k * 6 + (s)) * n + ti]
    const LetfBootstrapPool* pa = &b->pool_a[regime];
    const LetfBootstrapPool* pb = &b->pool_b[regime];
    /* an empty pool is served by a synthetic block (bootstrap.py:214-222, 273-300): block_size Student-t draws times a
     * regime volatility for SPY, QQQ = 1.25 SPY, TLT = -0.25 SPY; no momentum coin, no block index, same sub-section rule */
    const int syn_a = pa->n_blocks <= 0, syn_b = pb->n_blocks <= 0;
    if ((syn_a || syn_b) && (!tp->boot_syn || tp->boot_block_size != b->block_size)) return LETF_ERR_INVALID;
    const double syn_std = regime == 0 ? 0.011 : 0.035;
#define SYN(q, j) (tp->boot_syn[(((int64_t)k * 2 + (q)) * tp->boot_block_size + (j)) * n + ti] * syn_std)
    int row_a = syn_a ? 0 : pa->start[boot_pick(pa, have_last, last_a, bias, BU(0), BU(1), 0, 0.0)];
    if (len < b->block_size) row_a += (int)floor(BU(2) * (double)(b->block_size - len + 1));
    double prod_a = 1.0;
    for (int j = 0; j < len; j++) prod_a *= 1 + (syn_a ? SYN(0, row_a + j) : b->hist_spy[row_a + j]);
    const double spy_block_return = prod_a - 1;
    int row_b = syn_b ? 0 : pb->start[boot_pick(pb, have_last, last_b, bias, BU(3), BU(4), 1, spy_block_return)];
    if (len < b->block_size) row_b += (int)floor(BU(5) * (double)(b->block_size - len + 1));
#undef BU
    double prod_b = 1.0;
    for (int j = 0; j < len; j++) {
      spy[cur + j] = syn_a ? SYN(0, row_a + j) : b->hist_spy[row_a + j];
      qqq[cur + j] = syn_b ? SYN(1, row_b + j) * 1.25 : b->hist_qqq[row_b + j];
      tlt[cur + j] = syn_b ? -SYN(1, row_b + j) * 0.25 : b->hist_tlt[row_b + j];
      prod_b *= 1 + qqq[cur + j];
    }
#undef SYN
    last_a = spy_block_return; last_b = prod_b - 1; have_last = 1;
    cur += len; k++;
  }
  /* shared correlated Student-t noise, mean-preserving blend (bootstrap.py:386-449) */
  if (b->sampler_weight < 1.0 && G > 0) {
    const double nw = 1.0 - b->sampler_weight;
    const double m0 = np_pairwise_sum(spy, G) / (double)G, m1 = np_pairwise_sum(qqq, G) / (double)G,
                 m2 = np_pairwise_sum(tlt, G) / (double)G;
    for (int t = 0; t < G; t++) {
      const int rg = regf[t];
      const double s0 = rg == 0 ? 0.007 : 0.022, s1 = s0 * 1.35, s2 = s0 * 0.5;
      const double (*C)[3] = b->sampler_chol[rg];
      const double z0 = tp->boot_t[((int64_t)t * 3 + 0) * n + ti], z1 = tp->boot_t[((int64_t)t * 3 + 1) * n + ti],
                   z2 = tp->boot_t[((int64_t)t * 3 + 2) * n + ti];
      const double c0 = C[0][0] * z0 + C[0][1] * z1 + C[0][2] * z2;
      const double c1 = C[1][0] * z0 + C[1][1] * z1 + C[1][2] * z2;
      const double c2 = C[2][0] * z0 + C[2][1] * z1 + C[2][2] * z2;
      spy[t] = b->sampler_weight * spy[t] + nw * (c0 * s0 + m0);
      qqq[t] = b->sampler_weight * qqq[t] + nw * (c1 * s1 + m1);
      tlt[t] = b->sampler_weight * tlt[t] + nw * (c2 * s2 + m2);
    }
  }
  /* second blend, engine.py:57-79 */
  {
    const double W = b->engine_weight;
    const double noise_w = clipd(1.0 - W, 0.0, 1.0);
    for (int t = 0; t < G; t++) {
      const int rg = regf[t];
      const double (*C)[3] = b->engine_chol[rg];
      const double z0 = tp->x[((int64_t)t * 3 + 0) * n + ti], z1 = tp->x[((int64_t)t * 3 + 1) * n + ti],
                   z2 = tp->x[((int64_t)t * 3 + 2) * n + ti];
      const double x0 = C[0][0] * z0 + C[0][1] * z1 + C[0][2] * z2;
      const double x1 = C[1][0] * z0 + C[1][1] * z1 + C[1][2] * z2;
      const double x2 = C[2][0] * z0 + C[2][1] * z1 + C[2][2] * z2;
      const double sd = p->regime_daily_std[rg], qsd = 1.35 * sd, tsd = 0.55 * sd;
      const double n0 = noise_w * sd * 0.55 * x0, n1 = noise_w * qsd * 0.45 * x1, n2 = noise_w * tsd * 0.35 * x2;
      spy[t] = W * spy[t] + (1.0 - W) * (spy[t] + n0);
      qqq[t] = W * qqq[t] + (1.0 - W) * (qqq[t] + n1);
      tlt[t] = W * tlt[t] + (1.0 - W) * (tlt[t] + n2);
    }
  }
  return 0;
}

/* one trial.  Returns 0, or LETF_ERR_NONFINITE when the reference would raise. */
static int oracle_trial(const LetfParams* p, const LetfTape* tp, int64_t k, int64_t col,
                        const LetfOutputs* out, LetfOracleSeries* dump, Journal* full_journal,
                        int journal_strategy) {
  const int T = p->sim_days;
  const int64_t n = tp->n_tapes;
  const int64_t ti = p->antithetic_pairs ? (k >> 1) : k;
  const int anti = p->antithetic_pairs ? (int)(k & 1) : 0;
  const int buffer = p->buffer_days;
  const int G = T + buffer;
  const int so = p->random_start ? tp->start_offset[ti] : 0;
  uint32_t flags = 0;
  if (ti < 0 || ti >= n) return LETF_ERR_INVALID;

  int* regf = (int*)calloc(G > 0 ? G : 1, sizeof(int));
  double* gen = (double*)calloc((size_t)(5 * (G > 0 ? G : 1)), sizeof(double));
  double *spy = gen, *qqq = gen + G, *tlt = gen + 2 * G, *vixf = gen + 3 * G, *irxf = gen + 4 * G;

  /* ---- regime path, engine.py:255-302 (2 regimes; next regime is the only off-diagonal state) */
  if (p->market_mode != LETF_MARKET_NORMAL) {
    int cur = (tp->start_regime && tp->start_regime[ti]) ? 1 : 0, t = 0, ks = 0;   /* random_start.py:31 / :106 */
    while (t < G) {
      int spell;
      if (ks >= tp->max_spells) { flags |= LETF_FLAG_SPELL_OVERFLOW; spell = G; }
      else spell = tp->spells[(int64_t)ks * n + ti];
      ks++;
      if (spell < 1) spell = 1;
      int end = t + spell < G ? t + spell : G;
      for (int j = t; j < end; j++) regf[j] = cur;
      t = end;
      if (t < G) {
        double off = p->transition[cur][1 - cur];
        if (off > 0) cur = 1 - cur;
      }
    }
  }

  /* ---- market generator, engine.py:30-125 */
#define X(g, c) tp->x[((int64_t)(g) * 3 + (c)) * n + ti]
  if (p->market_mode == LETF_MARKET_JOINT_T) {
    /* calibration.py:480-590 */
    double h[3], Q[3][3], prev_z[3] = {0, 0, 0};
    int prev_regime = regf[0];
    for (int t = 0; t < G; t++) {
      int rg = regf[t];
      const LetfJointRegime* jp = &p->joint[rg];
      double nu = jp->nu, alpha = jp->garch_alpha, beta = jp->garch_beta;
      double tvs = nu > 2.0 ? (nu - 2.0) / nu : 0.5;
      double lv[3];
      for (int a = 0; a < 3; a++) { lv[a] = jp->cov[a][a] * tvs; if (lv[a] < 1e-8) lv[a] = 1e-8; }
      double ea = nu > 2.0 ? alpha * nu / (nu - 2.0) : alpha * 2.0;
      if (ea + beta >= 1.0) beta = dmax(0.70, 0.98 - ea);
      if (t == 0) {
        for (int a = 0; a < 3; a++) { h[a] = lv[a]; for (int b = 0; b < 3; b++) Q[a][b] = jp->rbar[a][b]; }
      } else if (rg != prev_regime) {
        double blend = 0.10;
        for (int a = 0; a < 3; a++) {
          h[a] = blend * lv[a] + (1 - blend) * h[a];
          for (int b = 0; b < 3; b++) Q[a][b] = blend * jp->rbar[a][b] + (1 - blend) * Q[a][b];
        }
      } else {
        for (int a = 0; a < 3; a++) {
          h[a] = (1 - alpha - beta) * lv[a] + alpha * (prev_z[a] * prev_z[a]) * h[a] + beta * h[a];
          if (h[a] < 1e-10) h[a] = 1e-10;
        }
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++)
            Q[a][b] = (1 - jp->dcc_a - jp->dcc_b) * jp->rbar[a][b] + jp->dcc_a * (prev_z[a] * prev_z[b]) + jp->dcc_b * Q[a][b];
      }
      double d[3], sh[3], cov_t[3][3], L[3][3];
      for (int a = 0; a < 3; a++) { double q = Q[a][a]; if (q < 1e-12) q = 1e-12; d[a] = sqrt(q); sh[a] = sqrt(h[a]); }
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) cov_t[a][b] = (sh[a] * sh[b]) * (Q[a][b] / (d[a] * d[b]));
      /* z ~ N(0, cov_t): Cholesky factor applied to the tape's standard normals.
       * (numpy's multivariate_normal factors with SVD; the tape harness uses this same
       * Cholesky so both sides share a factor - documented deviation) */
      chol3(cov_t, L);
      double x0 = X(t, 0), x1 = X(t, 1), x2 = X(t, 2), z[3];
      z[0] = L[0][0] * x0;
      z[1] = L[1][0] * x0 + L[1][1] * x1;
      z[2] = L[2][0] * x0 + L[2][1] * x1 + L[2][2] * x2;
      if (anti) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
      double u = tp->chi2[(int64_t)t * n + ti];
      double scale = sqrt(nu / dmax(u, 1e-12));
      double* outp[3] = {spy, qqq, tlt};
      for (int a = 0; a < 3; a++) {
        double xx = jp->mu[a] + z[a] * scale;
        double xc = clipd(xx, -0.95, 4.0);
        double hc = h[a] < 1e-10 ? 1e-10 : h[a];
        prev_z[a] = (xc - jp->mu[a]) / sqrt(hc);
        outp[a][t] = xc;
      }
      prev_regime = rg;
    }
    if (p->moment_matching && G > 0) {
      double* outp[3] = {spy, qqq, tlt};
      for (int a = 0; a < 3; a++) {
        double theo = (p->joint[0].mu[a] + p->joint[1].mu[a]) / 2.0;
        double act = np_pairwise_sum(outp[a], G) / (double)G;
        double sh = (theo - act) * 0.1;
        for (int t = 0; t < G; t++) outp[a][t] = outp[a][t] + sh;
      }
    }
  } else if (p->market_mode == LETF_MARKET_BOOTSTRAP) {
    int rc = boot_generate(p, tp, ti, n, regf, G, spy, qqq, tlt);
    if (rc) { free(regf); free(gen); return rc; }
  } else if (p->market_mode == LETF_MARKET_PARAM_T) {
    /* engine.py:81-97 */
    for (int t = 0; t < G; t++) {
      int rg = regf[t];
      double z0 = X(t, 0), z1 = X(t, 1), z2 = X(t, 2);
      const double (*C)[3] = p->param_t_chol[rg];
      double xa = C[0][0] * z0 + C[0][1] * z1 + C[0][2] * z2;
      double xb = C[1][0] * z0 + C[1][1] * z1 + C[1][2] * z2;
      double xc = C[2][0] * z0 + C[2][1] * z1 + C[2][2] * z2;
      spy[t] = p->regime_daily_mean[rg] + p->regime_daily_std[rg] * xa;
      qqq[t] = 1.15 * spy[t] + 0.006 * xb;
      tlt[t] = -0.12 * spy[t] + 0.004 * xc;
    }
  } else {
    for (int t = 0; t < G; t++) {
      double z0 = X(t, 0), z1 = X(t, 1), z2 = X(t, 2);
      const double (*C)[3] = p->normal_chol;
      spy[t] = p->normal_mu[0] + C[0][0] * z0;
      qqq[t] = p->normal_mu[1] + (C[1][0] * z0 + C[1][1] * z1);
      tlt[t] = p->normal_mu[2] + (C[2][0] * z0 + C[2][1] * z1 + C[2][2] * z2);
    }
  }
#undef X

  /* ---- VIX, engine.py:100-115; IRX, engine.py:117-123 */
  if (p->has_vix_dynamics) {
    if (G > 0) vixf[0] = p->vix_base[regf[0]];
    for (int t = 1; t < G; t++) {
      int rg = regf[t];
      const LetfVixRegime* v = &p->vix[rg];
      double denom = dmax(p->regime_daily_std[rg], 1e-4);
      double shock = dmax(-spy[t], 0.0) / denom;
      double vj = v->jump_scale * dmax(0.0, shock - v->jump_threshold);
      double cand = v->phi * vixf[t - 1] + (1 - v->phi) * v->target + vj + (0.0 + v->noise_std * tp->vix_n[(int64_t)t * n + ti]);
      vixf[t] = cand > 10.0 ? cand : 10.0; /* python max(10.0, x): x if x > 10.0 else 10.0 (NaN -> 10) */
    }
    for (int t = 0; t < G; t++)
      irxf[t] = clipd(p->irx_base[regf[t]] + (0.0 + 0.5 * tp->irx_n[(int64_t)t * n + ti]), 0.0, 15.0);
  } else {
    for (int t = 0; t < G; t++) { vixf[t] = p->const_vix; irxf[t] = p->const_rf * 100.0; }
  }

  if (dump) {
    if (dump->regime_full) for (int t = 0; t < G; t++) dump->regime_full[t] = regf[t];
    for (int c = 0; c < 5; c++) if (dump->gen[c]) memcpy(dump->gen[c], gen + (size_t)c * G, sizeof(double) * G);
  }

  /* ---- slice the window, engine.py:466-500 */
  size_t nT = (size_t)(T > 0 ? T : 1);
  double* buf = (double*)calloc(nT * (size_t)(10 + 2 * LETF_N_ASSETS + 16), sizeof(double));
  double *wspy = buf, *wqqq = buf + nT, *wtlt = buf + 2 * nT, *vix = buf + 3 * nT, *rf = buf + 4 * nT,
         *liq = buf + 5 * nT, *cred = buf + 6 * nT, *jump = buf + 7 * nT, *cash = buf + 8 * nT,
         *tltp = buf + 9 * nT;
  double* aret[LETF_N_ASSETS]; double* apr[LETF_N_ASSETS];
  for (int a = 0; a < LETF_N_ASSETS; a++) { aret[a] = buf + (10 + a) * nT; apr[a] = buf + (10 + LETF_N_ASSETS + a) * nT; }
  double* wk = buf + (10 + 2 * LETF_N_ASSETS) * nT; /* 16*T scratch */
  int* reg = (int*)calloc(nT, sizeof(int));
  for (int i = 0; i < T; i++) {
    reg[i] = regf[so + i]; wspy[i] = spy[so + i]; wqqq[i] = qqq[so + i]; wtlt[i] = tlt[so + i];
    vix[i] = vixf[so + i];
    if (!isfinite(wspy[i]) || !isfinite(wqqq[i]) || !isfinite(wtlt[i])) flags |= LETF_FLAG_NONFINITE_RET;
    if (!isfinite(vix[i]) || vix[i] <= 0) flags |= LETF_FLAG_BAD_VIX;
  }

  /* ---- stress + jump overlay, calibration.py:787-830, engine.py:502-512 */
  int jump_days = 0;
  if (p->has_stress && T > 0) {
    const LetfStressRegime* s0 = &p->stress[reg[0]];
    liq[0] = s0->liq_mu; cred[0] = s0->credit_mu; jump[0] = 0.0;
    for (int t = 1; t < T; t++) {
      const LetfStressRegime* s = &p->stress[reg[t]];
      double l = s->liq_mu + s->liq_phi * (liq[t - 1] - s->liq_mu) + (0.0 + s->liq_sigma * tp->liq_n[(int64_t)t * n + ti]);
      double c = s->credit_mu + s->credit_phi * (cred[t - 1] - s->credit_mu) + (0.0 + s->credit_sigma * tp->cred_n[(int64_t)t * n + ti]);
      liq[t] = clipd(l, 0.0, 3.0);
      cred[t] = clipd(c, 0.0, 3.0);
      double vamp = dmax((vix[t] - 25.0) / 30.0, 0.0);
      double jp = clipd(s->jump_base_prob + 0.008 * vamp + 0.003 * liq[t], 0.0, 0.01);
      jump[t] = 0.0;
      if (tp->jump_u[(int64_t)t * n + ti] < jp) {
        jump[t] = fabs(tp->jump_t[(int64_t)t * n + ti]) * s->jump_scale;
      }
      if (jump[t] > 0) jump_days++;
    }
    for (int t = 0; t < T; t++) {
      double sj = jump[t] * tp->sign[(int64_t)t * n + ti];
      wspy[t] = clipd(wspy[t] - sj, -0.95, 3.0);
      wqqq[t] = clipd(wqqq[t] - 1.12 * sj, -0.95, 4.0);
    }
  }
  for (int t = 0; t < T; t++)
    rf[t] = p->has_vix_dynamics ? clipd(irxf[so + t], 0.0, 20.0) / 100.0 : p->const_rf;

  /* ---- asset layers, engine.py:360-420 */
  for (int a = 0; a < LETF_N_ASSETS; a++) {
    if (!((p->asset_mask >> a) & 1)) continue;
    const LetfAssetParams* ap = &p->assets[a];
    const double* ubase = ap->underlying == 1 ? wqqq : (ap->underlying == 2 ? wtlt : wspy);
    double te_prev = 0.0, bias = 0.0;
    const double alpha = 2.0 / (126 + 1);
    const double arw = sqrt(dmax(1 - ap->te_rho * ap->te_rho, 1e-6));
    for (int t = 0; t < T; t++) {
      double u = ap->beta_applies ? ubase[t] * ap->beta : ubase[t];
      /* financing, engine.py:320-349 + calibration.py:833-857 + utils.py:169-216 */
      double spread;
      if (p->institutional && p->has_funding) {
        double irx_level = rf[t] * 100;
        double term = p->has_stress ? 1.25 - 1.1 * clipd(cred[t], 0.0, 3.0) : (vix[t] > 30 ? -0.25 : 1.10);
        double tnx = irx_level + term;
        double stress_v = dmax(vix[t] - 20.0, 0.0);
        double inv = dmax(-(tnx - irx_level), 0.0);
        spread = p->funding.base + p->funding.beta_vix * stress_v + p->funding.beta_inv_curve * inv;
        if (p->has_stress)
          spread += p->funding.beta_liquidity * clipd(liq[t], 0, 3) + p->funding.beta_credit * clipd(cred[t], 0, 3);
        spread = clipd(spread, p->funding.min_spread, p->funding.max_spread);
      } else spread = ap->borrow_spread;
      double borrow = 0.0;
      double bl = ap->leverage - 1.0;
      if (bl > 0) { double rate = dmax(rf[t] + spread, 0.0); borrow = bl * rate / 252; }
      double lev = ap->leverage * u - ap->expense_ratio / 252 - borrow;
      /* tracking error, engine.py:157-203 */
      double te_raw = 0.0;
      if (t >= 1) {
        double vm = clipd(sqrt(vix[t] / 20.0), 0.5, 1.5);
        double rm = reg[t] == 0 ? 1.0 : 1.15;
        double innov = tp->te_t[((int64_t)a * T + t) * n + ti];
        innov *= ap->te_base * vm * rm * 1.0;
        double ds = u < 0 ? ap->te_downside : 0.98;
        double mm = (1.0 + 0.3 * fabs(u)) * ds;
        te_raw = ap->te_rho * te_prev + arw * innov * mm;
      }
      te_prev = te_raw;
      bias = alpha * te_raw + (1 - alpha) * bias;
      double te = te_raw - bias;
      if (ap->te_clip > 0) te = ap->te_clip * tanh(te / ap->te_clip);
      aret[a][t] = clipd(lev + te, -0.999, 10.0);
    }
  }

  /* ---- frame, engine.py:537-550 */
  int bad = 0;
  double vmin = INFINITY, vmax = -INFINITY;
  for (int t = 0; t < T; t++) {
    cash[t] = p->cash_rate[reg[t]] / 252;
    if (vix[t] < vmin) vmin = vix[t];
    if (vix[t] > vmax) vmax = vix[t];
  }
  for (int a = 0; a < LETF_N_ASSETS; a++) {
    if (!((p->asset_mask >> a) & 1)) continue;
    double cp = 1.0;
    for (int t = 0; t < T; t++) {
      cp = t == 0 ? (1 + aret[a][0]) : cp * (1 + aret[a][t]);
      apr[a][t] = cp * 100;
      if (!isfinite(aret[a][t]) || !isfinite(apr[a][t]) || !(apr[a][t] > 0.0)) bad = 1;
    }
  }
  {
    double cp = 1.0;
    for (int t = 0; t < T; t++) {
      cp = t == 0 ? (1 + wtlt[0]) : cp * (1 + wtlt[t]);
      tltp[t] = cp * 100;
      if (!isfinite(wtlt[t]) || !isfinite(tltp[t]) || !(tltp[t] > 0.0)) bad = 1;
    }
  }
  /* validate_simulation_layers, engine.py:206-231: the repair pass re-applies clips the series
   * already satisfy, so it only sets the flag */
  if (T == 0 || bad || !(vmin >= 5.0) || !(vmax <= 120.0)) flags |= LETF_FLAG_REPAIRED;

  if (dump) {
    for (int a = 0; a < LETF_N_ASSETS; a++) if (dump->asset_ret[a]) memcpy(dump->asset_ret[a], aret[a], sizeof(double) * T);
    if (dump->liq) memcpy(dump->liq, liq, sizeof(double) * T);
    if (dump->cred) memcpy(dump->cred, cred, sizeof(double) * T);
    if (dump->jump) memcpy(dump->jump, jump, sizeof(double) * T);
  }

  /* ---- strategies, engine.py:569-617 */
  Frame f;
  f.T = T; f.cash = cash; f.tlt_ret = wtlt; f.tlt_price = tltp; f.vix = vix; f.regime = reg;
  for (int a = 0; a < LETF_N_ASSETS; a++) { f.ret[a] = aret[a]; f.price[a] = apr[a]; }
  double* ec = (double*)calloc(nT, sizeof(double));
  const int64_t N = out->n_trials;
  for (int s = 0; s < p->n_strategies; s++) {
    const LetfStrategyParams* sp = &p->strategies[s];
    Journal jr; memset(&jr, 0, sizeof(jr));
    Journal* jp = sp->taxable ? &jr : NULL;
    if (jp && full_journal && journal_strategy == s) { jr = *full_journal; jr.count = 0; jr.volume = 0; jr.n = 0; memset(jr.pos, 0, sizeof(jr.pos)); }
    int nt = run_strategy(p, sp, &f, ec, jp, wk);
    double fw, mdd;
    max_drawdown_and_final(ec, T, &fw, &mdd);
    out->final_wealth[(int64_t)s * N + col] = fw;
    out->max_dd[(int64_t)s * N + col] = mdd;
    out->num_trades[(int64_t)s * N + col] = nt;
    if (out->journal_count) out->journal_count[(int64_t)s * N + col] = jp ? (int32_t)jr.count : 0;
    if (out->journal_volume) out->journal_volume[(int64_t)s * N + col] = jp ? jr.volume : 0.0;
    if (jp && full_journal && journal_strategy == s) { full_journal->n = jr.n; full_journal->count = jr.count; full_journal->volume = jr.volume; }
    if (dump && dump->equity) memcpy(dump->equity + (size_t)s * T, ec, sizeof(double) * T);
  }
  if (out->start_offset) out->start_offset[col] = so;
  if (out->flags) out->flags[col] = flags;
  if (out->regime_days) {
    int c0 = 0, c1 = 0;
    for (int t = 0; t < T; t++) { if (reg[t] == 0) c0++; else c1++; }
    out->regime_days[col] = c0; out->regime_days[N + col] = c1;
  }
  if (out->jump_days) out->jump_days[col] = jump_days;

  free(ec); free(reg); free(buf); free(gen); free(regf);
  return (flags & (LETF_FLAG_NONFINITE_RET | LETF_FLAG_BAD_VIX)) ? LETF_ERR_NONFINITE : 0;
}

/* Public: trials [first, first+n) -> columns [0,n) of out (host pointers).  Returns the number
 * of trials for which the reference would have raised (non-finite), or <0 on error. */
int letf_oracle_run(const LetfParams* p, const LetfTape* tape, int64_t first, int64_t n_trials,
                    const LetfOutputs* out) {
  if (!p || !tape || !out || p->abi_version != LETF_ABI_VERSION) return -1;
  int bad = 0;
  for (int64_t j = 0; j < n_trials; j++) {
    int rc = oracle_trial(p, tape, first + j, j, out, NULL, NULL, -1);
    if (rc == LETF_ERR_INVALID) return -2;
    if (rc) bad++;
  }
  return bad;
}

/* one trial with per-day series dumped */
int letf_oracle_run_series(const LetfParams* p, const LetfTape* tape, int64_t trial,
                           const LetfOutputs* out, LetfOracleSeries* series) {
  if (!p || !tape || !out) return -1;
  return oracle_trial(p, tape, trial, 0, out, series, NULL, -1);
}

/* one trial, one taxable strategy, with the full Trade_List (trade.py:83-91).
 * action: +1 BUY, -1 SELL.  Returns number of journal entries written (<= cap). */
int letf_oracle_trade_list(const LetfParams* p, const LetfTape* tape, int64_t trial, int strategy,
                           const LetfOutputs* out, int cap, int32_t* day, int32_t* asset,
                           int32_t* action, double* shares, double* price) {
  Journal j; memset(&j, 0, sizeof(j));
  j.cap = cap; j.day = day; j.asset = asset; j.action = action; j.shares = shares; j.price = price;
  int rc = oracle_trial(p, tape, trial, 0, out, NULL, &j, strategy);
  if (rc == LETF_ERR_INVALID) return -1;
  return j.n;
}

/* run_strategy_fixed on caller-supplied series (host arrays [T][n], day-major) */
int letf_oracle_run_strategies(const LetfParams* p, const LetfSeries* sr, const LetfOutputs* out) {
  const int T = sr->sim_days;
  const int64_t n = sr->n_paths, N = out->n_trials;
  size_t nT = (size_t)(T > 0 ? T : 1);
  double* buf = (double*)calloc(nT * (size_t)(4 + 2 * LETF_N_ASSETS + 16), sizeof(double));
  int* reg = (int*)calloc(nT, sizeof(int));
  double* ec = (double*)calloc(nT, sizeof(double));
  for (int64_t k = 0; k < n; k++) {
    Frame f; f.T = T;
    f.cash = buf; f.tlt_ret = buf + nT; f.tlt_price = buf + 2 * nT; f.vix = buf + 3 * nT; f.regime = reg;
    for (int a = 0; a < LETF_N_ASSETS; a++) { f.ret[a] = buf + (4 + a) * nT; f.price[a] = buf + (4 + LETF_N_ASSETS + a) * nT; }
    double* wk = buf + (4 + 2 * LETF_N_ASSETS) * nT;
    for (int t = 0; t < T; t++) {
      reg[t] = sr->regime ? sr->regime[(int64_t)t * n + k] : 0;
      f.cash[t] = p->cash_rate[reg[t]] / 252;
      f.vix[t] = sr->vix ? sr->vix[(int64_t)t * n + k] : 15.0;
      f.tlt_ret[t] = sr->tlt_ret ? sr->tlt_ret[(int64_t)t * n + k] : 0.0;
    }
    for (int a = 0; a < LETF_N_ASSETS; a++) {
      double cp = 1.0;
      for (int t = 0; t < T; t++) {
        f.ret[a][t] = sr->asset_ret[a] ? sr->asset_ret[a][(int64_t)t * n + k] : 0.0;
        cp = t == 0 ? (1 + f.ret[a][0]) : cp * (1 + f.ret[a][t]);
        f.price[a][t] = cp * 100;
      }
    }
    for (int s = 0; s < p->n_strategies; s++) {
      const LetfStrategyParams* sp = &p->strategies[s];
      Journal jr; memset(&jr, 0, sizeof(jr));
      int nt = run_strategy(p, sp, &f, ec, sp->taxable ? &jr : NULL, wk);
      double fw, mdd;
      max_drawdown_and_final(ec, T, &fw, &mdd);
      out->final_wealth[(int64_t)s * N + k] = fw;
      out->max_dd[(int64_t)s * N + k] = mdd;
      out->num_trades[(int64_t)s * N + k] = nt;
      if (out->journal_count) out->journal_count[(int64_t)s * N + k] = sp->taxable ? (int32_t)jr.count : 0;
      if (out->journal_volume) out->journal_volume[(int64_t)s * N + k] = sp->taxable ? jr.volume : 0.0;
    }
  }
  free(ec); free(reg); free(buf);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Stand-alone tape filler for the CPU-baseline timing leg (bench.py): draws every variate the
 * path consumes with a conventional CPU generator (xoshiro256++, polar normals,
 * Marsaglia-Tsang gamma), i.e. the work numpy's Generator does for the reference. */
typedef struct { uint64_t s[4]; int has; double spare; } Xo;
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t xo_next(Xo* g) {
  uint64_t* s = g->s; uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return r;
}
static void xo_seed(Xo* g, uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
  for (int i = 0; i < 4; i++) { z += 0x9E3779B97F4A7C15ULL; uint64_t x = z; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; g->s[i] = x ^ (x >> 31); }
  g->has = 0;
}
static double xo_u(Xo* g) { return (double)(xo_next(g) >> 11) * (1.0 / 9007199254740992.0); }
static double xo_n(Xo* g) {
  if (g->has) { g->has = 0; return g->spare; }
  double a, b, s;
  do { a = 2 * xo_u(g) - 1; b = 2 * xo_u(g) - 1; s = a * a + b * b; } while (s >= 1 || s == 0);
  double m = sqrt(-2 * log(s) / s);
  g->spare = b * m; g->has = 1;
  return a * m;
}
static double xo_gamma(Xo* g, double a) {
  if (a < 1) { double u = xo_u(g); return xo_gamma(g, a + 1) * pow(u, 1 / a); }
  double d = a - 1.0 / 3, c = 1 / sqrt(9 * d);
  for (;;) {
    double x, v;
    do { x = xo_n(g); v = 1 + c * x; } while (v <= 0);
    v = v * v * v;
    double u = xo_u(g);
    if (u < 1 - 0.0331 * x * x * x * x) return d * v;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) return d * v;
  }
}
static double xo_t(Xo* g, double df) { double nn = xo_n(g); return nn / sqrt(2 * xo_gamma(g, df / 2) / df); }

/* fills tape columns [0, n_tapes) (host arrays, already allocated) */
int letf_oracle_fill_tape(const LetfParams* p, uint64_t seed, LetfTape* tp) {
  const int T = tp->sim_days, G = tp->gen_days;
  const int64_t n = tp->n_tapes;
  for (int64_t c = 0; c < n; c++) {
    Xo g; xo_seed(&g, seed * 0x100000001B3ULL + (uint64_t)c);
    int buf = p->buffer_days;
    int mn = buf / 4 < 50 ? buf / 4 : 50;
    tp->start_offset[c] = (p->random_start && buf > 0) ? mn + (int)(xo_u(&g) * (buf - mn)) : 0;
    /* spells + regime path (needed to pick nu per day) */
    int* regf = (int*)calloc(G > 0 ? G : 1, sizeof(int));
    int cur = 0, t = 0, ks = 0;
    if (tp->start_regime) { cur = (p->start_regime_prob1 > 0 && xo_u(&g) < p->start_regime_prob1) ? 1 : 0; tp->start_regime[c] = cur; }
    while (ks < tp->max_spells) {
      int spell;
      if (p->n_duration[cur] > 0) spell = p->duration_samples[cur][(int)(xo_u(&g) * p->n_duration[cur])];
      else {
        double ps = clipd(p->transition[cur][cur], 0.80, 0.995);
        spell = (int)ceil(log(1.0 - xo_u(&g)) / log(ps));
      }
      if (spell < 1) spell = 1;
      tp->spells[(int64_t)ks * n + c] = spell;
      ks++;
      if (t < G) {
        int end = t + spell < G ? t + spell : G;
        for (int j = t; j < end; j++) regf[j] = cur;
        t = end;
        if (p->transition[cur][1 - cur] > 0) cur = 1 - cur;
      } else if (ks > 8) { for (; ks < tp->max_spells; ks++) tp->spells[(int64_t)ks * n + c] = G; break; }
    }
    for (int gday = 0; gday < G; gday++) {
      for (int a = 0; a < 3; a++)
        tp->x[((int64_t)gday * 3 + a) * n + c] = (p->market_mode == LETF_MARKET_PARAM_T || p->market_mode == LETF_MARKET_BOOTSTRAP)
                                                     ? xo_t(&g, p->param_t_df) : xo_n(&g);
      if (p->market_mode == LETF_MARKET_BOOTSTRAP && tp->boot_t)
        for (int a = 0; a < 3; a++) tp->boot_t[((int64_t)gday * 3 + a) * n + c] = xo_t(&g, p->param_t_df);
      if (p->market_mode == LETF_MARKET_JOINT_T) tp->chi2[(int64_t)gday * n + c] = 2 * xo_gamma(&g, p->joint[regf[gday]].nu / 2);
      if (p->has_vix_dynamics) { tp->vix_n[(int64_t)gday * n + c] = xo_n(&g); tp->irx_n[(int64_t)gday * n + c] = xo_n(&g); }
    }
    if (p->market_mode == LETF_MARKET_BOOTSTRAP && tp->boot_len && tp->boot_u)
      for (int kb = 0; kb < tp->max_blocks; kb++) {
        tp->boot_len[(int64_t)kb * n + c] = (int)ceil(log(1.0 - xo_u(&g)) / log(1.0 - 1.0 / p->boot.block_mean));
        if (tp->boot_len[(int64_t)kb * n + c] < 1) tp->boot_len[(int64_t)kb * n + c] = 1;
        for (int q = 0; q < 6; q++) tp->boot_u[((int64_t)kb * 6 + q) * n + c] = xo_u(&g);
        if (tp->boot_syn)
          for (int q = 0; q < 2 * tp->boot_block_size; q++)
            tp->boot_syn[((int64_t)kb * 2 * tp->boot_block_size + q) * n + c] = xo_t(&g, p->param_t_df);
      }
    for (int i = 0; i < T; i++) {
      if (p->has_stress) {
        tp->liq_n[(int64_t)i * n + c] = xo_n(&g); tp->cred_n[(int64_t)i * n + c] = xo_n(&g);
        tp->jump_u[(int64_t)i * n + c] = xo_u(&g); tp->jump_t[(int64_t)i * n + c] = xo_t(&g, 5.0);
        tp->sign[(int64_t)i * n + c] = xo_u(&g) < 0.5 ? -1.0 : 1.0;
      }
      for (int a = 0; a < LETF_N_ASSETS; a++)
        if ((p->asset_mask >> a) & 1) tp->te_t[((int64_t)a * T + i) * n + c] = xo_t(&g, p->assets[a].te_df);
    }
    free(regf);
  }
  return 0;
}

int letf_oracle_abi_version(void) { return LETF_ABI_VERSION; }
