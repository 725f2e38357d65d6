// letf_rolling.cuh — streaming (one value in, one value out) forms of the pandas rolling
// statistics the reference's strategy rules are written with (letf/strategy.py).
//
// EXACT = true reproduces pandas' own online algorithms (removal before addition, Kahan
// compensated sums, Welford variance with compensated mean update, the
// "all values in the window identical" short-cuts), so that fp64 parity-mode signals agree
// with the reference to the last bits and trade decisions are identical.
// EXACT = false (fp32 throughput mode) drops the compensation terms.
#pragma once
#include "letf_math.cuh"

namespace letf {

template <typename R> LETF_HD R nan_() { return (R)NAN; }

// rolling sum / mean ------------------------------------------------------------------------
template <typename R, bool EXACT> struct RollSum {
  R s, ca, cr, prev;
  int nobs, ncs, neg;
  LETF_HD void init() { s = ca = cr = (R)0; prev = nan_<R>(); nobs = ncs = neg = 0; }
  LETF_HD void remove(R v) {
    nobs--;
    if (EXACT) { R y = -v - cr, t = s + y; cr = t - s - y; s = t; if (v < (R)0) neg--; }
    else s -= v;
  }
  LETF_HD void add(R v) {
    nobs++;
    if (EXACT) {
      R y = v - ca, t = s + y; ca = t - s - y; s = t;
      if (v < (R)0) neg++;
      ncs = (v == prev) ? ncs + 1 : 1;
      prev = v;
    } else s += v;
  }
  LETF_HD R sum(int mp) const {
    if (nobs < mp) return nan_<R>();
    if (EXACT && ncs >= nobs) return prev * (R)nobs;
    return s;
  }
  LETF_HD R mean(int mp) const {
    if (nobs < mp || nobs <= 0) return nan_<R>();
    R r = Mth<R>::div_(s, (R)nobs);
    if (EXACT) {
      if (ncs >= nobs) r = prev;
      else if (neg == 0 && r < (R)0) r = (R)0;
      else if (neg == nobs && r > (R)0) r = (R)0;
    }
    return r;
  }
};

// rolling variance (ddof = 1) -----------------------------------------------------------------
// pandas 3.0.2 roll_var: Welford add/remove with Kahan-compensated mean updates; when a removal
// collapses the sum of squared deviations (ssq_after < 1e3 * eps * ssq_before, e.g. the window
// became all-identical) pandas recomputes the window from scratch.  remove() reports that case
// and the caller re-adds the surviving window from its history ring (see Trial::revar).
template <typename R, bool EXACT> struct RollVar {
  R mean, ssq, ca, cr;
  int nobs;
  LETF_HD void init() { mean = ssq = ca = cr = (R)0; nobs = 0; }
  LETF_HD bool remove(R v) {
    const R before = ssq;
    nobs--;
    if (nobs) {
      if (EXACT) {
        R pm = mean - cr, y = v - cr, t = y - mean;
        cr = t + mean - y;
        mean = mean - t / (R)nobs;
        ssq = ssq - (v - pm) * (v - mean);
      } else {
        R pm = mean, t = v - mean;
        mean = mean - Mth<R>::div_(t, (R)nobs);
        ssq = ssq - (v - pm) * (v - mean);
      }
    } else { mean = (R)0; ssq = (R)0; }
    if (EXACT) return ssq < (R)(1e3 * 2.220446049250313e-16) * before;
    if (ssq < (R)0) ssq = (R)0;      // fp32: keep the state non-negative instead of recomputing
    return false;
  }
  LETF_HD void add(R v) {
    nobs++;
    if (EXACT) {
      R pm = mean - ca, y = v - ca, t = y - mean;
      ca = t + mean - y;
      mean = mean + t / (R)nobs;
      ssq = ssq + (v - pm) * (v - mean);
    } else {
      R pm = mean, t = v - mean;
      mean = mean + Mth<R>::div_(t, (R)nobs);
      ssq = ssq + (v - pm) * (v - mean);
    }
  }
  // rolling(w, min_periods = mp).std()
  LETF_HD R stdev(int mp) const {
    if (nobs < mp || nobs <= 1) return nan_<R>();
    R r = Mth<R>::div_(ssq, (R)(nobs - 1));
    if (r < (R)0) r = (R)0;
    return Mth<R>::sqrt_(r);
  }
};

// rolling skew (bias-corrected G1) + the window mean, from online central moments -----------------
// The moment state is kept in fp64 in BOTH precisions: the add/remove updates of m2 / m3 cancel
// catastrophically in fp32 (1-3 % wealth error on the skew-driven rules over 10 years, measured);
// it is ~25 fp64 operations per day and only in builds that carry a skew rule (S14, S17).
template <typename R, bool EXACT> struct RollSkew {
  double mean, m2, m3;
  R prev;
  int nobs, ncs;
  LETF_HD void init() { mean = m2 = m3 = 0.0; prev = nan_<R>(); nobs = ncs = 0; }
  LETF_HD void remove(R vr) {
    const double v = (double)vr;
    int nn = nobs - 1;
    if (nn == 0) { mean = m2 = m3 = 0.0; }
    else {
      double delta = v - mean, dn = delta / (double)nn, term1 = delta * dn * (double)nobs;
      m3 = m3 - (term1 * dn * (double)(nobs + 1) - 3.0 * dn * m2);
      mean = mean - dn;
      m2 = m2 - term1;
    }
    nobs = nn;
  }
  LETF_HD void add(R vr) {
    const double v = (double)vr;
    int nn = nobs + 1;
    double delta = v - mean, dn = delta / (double)nn, term1 = delta * dn * (double)nobs;
    m3 = m3 + term1 * dn * (double)(nn - 2) - 3.0 * dn * m2;
    m2 = m2 + term1;
    mean = mean + dn;
    nobs = nn;
    if (EXACT) { ncs = (vr == prev) ? ncs + 1 : 1; prev = vr; }
  }
  LETF_HD R skew(int mp) const {
    if (nobs < mp || nobs < 3) return nan_<R>();
    double d = (double)nobs, B = m2 / d;
    if (EXACT && ncs >= nobs) return (R)0;
    if (B <= 1e-14) return nan_<R>();
    return (R)(::sqrt(d * (d - 1.0)) / (d - 2.0) * (m3 / d) / (B * ::sqrt(B)));
  }
};

}  // namespace letf
