// Host-side hot loop of the sampled-trial tax path: realised gains / losses per tax year of ONE trade list, with wash-sale
// basis adjustments and lot selection (reference: letf/integration.py:113-226 over letf/tax/wash_sale.py:149-397 and
// letf/tax/lot_selection.py; Python restatement: letf_monte_carlo_b200/tax_engine.py `realised_by_year`, which stays as the
// readable reference and is compared with this port value for value in tests/test_tax_engine.py).
//
// Why native: create_summary_statistics taxes <= 50 trade lists per taxable strategy (reporting.py:457-487); at the benchmarked
// horizon a volatility rule journals 2 000 - 7 000 trades, and the per-trade Python bookkeeping of 650 such lists costs ~30 s
// next to a 23 ms Monte Carlo.  No GPU work here: plain C++ compiled into the same library.
//
// Arithmetic follows the Python statement operation for operation (same summation order, same stable sorts), so the two
// agree to the last bit on everything but library rounding.
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <map>
#include <utility>
#include <vector>
#include "../../include/letf_b200.h"

namespace {

constexpr double kEpsShares = 0.001;
constexpr int kDaysPerYear = 252;

struct Lot { double shares, basis; int start; };

// lot_selection.py: indices of `lots` to draw from, in order, until `need` shares are covered
static void lots_to_sell(const std::vector<Lot>& lots, double need, int method, int sale_day, double sale_price, std::vector<int>& out) {
  if (method == 0) {          // fifo (the default election): the book's own order - no candidate list to build
    out.clear();
    double left = need;
    for (int i = 0; i < (int)lots.size(); i++) {
      if (!(lots[i].shares > kEpsShares)) continue;
      if (left <= kEpsShares) break;
      out.push_back(i);
      left -= lots[i].shares;
    }
    return;
  }
  std::vector<int> live;
  for (int i = 0; i < (int)lots.size(); i++) if (lots[i].shares > kEpsShares) live.push_back(i);
  auto age = [&](int i) { return sale_day - lots[i].start; };
  std::vector<int> order;
  switch (method) {
    case 1: order.assign(live.rbegin(), live.rend()); break;                                             // lifo
    case 2: order = live; std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lots[a].basis > lots[b].basis; }); break;   // hifo
    case 3: order = live; std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lots[a].basis < lots[b].basis; }); break;   // lofo
    case 4: case 5: {                                                                                      // ltfo / stfo
      std::vector<int> lt, st;
      for (int i : live) (age(i) > 365 ? lt : st).push_back(i);
      if (method == 4) {
        auto desc = [&](int a, int b) { return age(a) > age(b); };
        std::stable_sort(lt.begin(), lt.end(), desc); std::stable_sort(st.begin(), st.end(), desc);
        order = lt; order.insert(order.end(), st.begin(), st.end());
      } else {
        auto asc = [&](int a, int b) { return age(a) < age(b); };
        std::stable_sort(lt.begin(), lt.end(), asc); std::stable_sort(st.begin(), st.end(), asc);
        order = st; order.insert(order.end(), lt.begin(), lt.end());
      }
      break;
    }
    case 6: case 7: {                                                                                      // mintax / specific
      auto impact = [&](int i) {
        const double g = sale_price - lots[i].basis;
        return g * (g >= 0 ? (age(i) > 365 ? 0.20 : 0.37) : 0.37);
      };
      order = live;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return impact(a) < impact(b); });
      break;
    }
    default: order = live;                                                                                 // fifo
  }
  out.clear();
  double left = need;
  for (int i : order) {
    if (left <= kEpsShares) break;
    out.push_back(i);
    left -= lots[i].shares;
  }
}

}  // namespace

// action: 1 = BUY, 0 = SELL, anything else is ignored.  asset: small non-negative ids.  out: [n_years][5] =
// st_gains, st_losses, lt_gains, lt_losses, wash_sale_disallowed per tax year (day / 252); trades of later years are dropped
// from the table (they cannot occur: the journal ends with the horizon) but still move the lots.
extern "C" int letf_tax_realised_by_year(const int32_t* day, const int32_t* asset, const int32_t* action, const double* shares,
                                         const double* price, int64_t n, int lot_method, int n_years, double* out) {
  if (n < 0 || n_years < 0 || (n > 0 && (!day || !asset || !action || !shares || !price)) || (n_years > 0 && !out)) return LETF_ERR_INVALID;
  if (lot_method < 0 || lot_method > 7) return LETF_ERR_INVALID;
  for (int64_t i = 0; i < (int64_t)n_years * 5; i++) out[i] = 0.0;
  int n_assets = 0;
  for (int64_t i = 0; i < n; i++) { if (asset[i] < 0) return LETF_ERR_INVALID; n_assets = std::max(n_assets, asset[i] + 1); }

  // ---- wash sales (tax/wash_sale.py:149-397): basis adjustment and holding-period start per (asset, replacement buy day)
  std::map<std::pair<int, int>, double> basis_adj;
  std::map<std::pair<int, int>, int> hold;
  std::vector<std::vector<int64_t>> by_asset(n_assets);
  for (int64_t i = 0; i < n; i++) by_asset[asset[i]].push_back(i);
  for (int a = 0; a < n_assets; a++) {
    std::vector<int64_t>& rows = by_asset[a];
    std::stable_sort(rows.begin(), rows.end(), [&](int64_t x, int64_t y) { return day[x] < day[y]; });     // same-day trades keep their order
    std::vector<int64_t> buys;
    for (int64_t r : rows) if (action[r] == 1) buys.push_back(r);
    std::vector<int> bdays(buys.size());
    for (size_t k = 0; k < buys.size(); k++) bdays[k] = day[buys[k]];
    // running sums over the buys that hold shares: FIFO cost of a sale over ALL buys (the reference never depletes them)
    std::vector<int64_t> live;
    std::vector<double> cum_sh, cum_cost;
    {
      double s = 0.0, c = 0.0;
      bool first = true;
      for (int64_t b : buys) {
        if (!(shares[b] > kEpsShares)) continue;
        live.push_back(b);
        // (itertools.accumulate: the first element as is, then running additions)
        if (first) { s = shares[b]; c = shares[b] * price[b]; first = false; }
        else { s = s + shares[b]; c = c + shares[b] * price[b]; }
        cum_sh.push_back(s); cum_cost.push_back(c);
      }
    }
    std::vector<char> used(buys.size(), 0);
    for (int64_t r : rows) {
      if (action[r] != 0) continue;
      const double sh = shares[r], px = price[r];
      const int d = day[r];
      double cost = 0.0;
      int first_day = -1;
      if (sh > kEpsShares && !live.empty()) {
        first_day = day[live[0]];
        const size_t m = std::lower_bound(cum_sh.begin(), cum_sh.end(), sh) - cum_sh.begin();
        if (m >= live.size()) cost = cum_cost.back();
        else {
          const double prev_sh = m ? cum_sh[m - 1] : 0.0;
          cost = m ? cum_cost[m - 1] : 0.0;
          const double left = sh - prev_sh;
          if (left > kEpsShares) cost += std::min(left, shares[live[m]]) * price[live[m]];
        }
      }
      const double loss = -(sh * px - cost);
      if (loss <= 0) continue;
      // replacement buy: the nearest one within 30 days (not the same day, not used yet); ties: the earlier buy
      const size_t lo = std::lower_bound(bdays.begin(), bdays.end(), d - 30) - bdays.begin();
      const size_t hi = std::upper_bound(bdays.begin(), bdays.end(), d + 30) - bdays.begin();
      int best = -1, best_dist = 0;
      for (size_t k = lo; k < hi; k++) {
        if (bdays[k] == d || used[k]) continue;
        const int dist = std::abs(bdays[k] - d);
        if (best < 0 || dist < best_dist) { best = (int)k; best_dist = dist; }
      }
      if (best < 0) continue;
      const int64_t rb = buys[best];
      const double rsh = shares[rb];
      const double disallowed = rsh >= sh ? loss : loss * (rsh / sh);
      basis_adj[{a, day[rb]}] += disallowed;
      hold[{a, day[rb]}] = first_day >= 0 ? first_day : d;
      used[best] = 1;
    }
  }

  // ---- realised gains by year (integration.py:113-226), trades in journal order
  std::vector<std::vector<int>> buy_days(n_assets);
  for (int64_t i = 0; i < n; i++) if (action[i] == 1) buy_days[asset[i]].push_back(day[i]);
  for (auto& v : buy_days) std::sort(v.begin(), v.end());
  std::vector<std::vector<Lot>> book(n_assets);
  std::vector<int> pick;
  for (int64_t i = 0; i < n; i++) {
    const int a = asset[i], d = day[i];
    const int year = d >= 0 ? d / kDaysPerYear : -1;
    double* y = (year >= 0 && year < n_years) ? out + (int64_t)year * 5 : nullptr;
    if (action[i] == 1) {
      const double sh = shares[i];
      double adj = 0.0;
      auto it = basis_adj.find({a, d});
      if (it != basis_adj.end()) adj = it->second;
      const double cost = sh * price[i] + adj;
      int start = 0;
      auto ih = hold.find({a, d});
      if (ih != hold.end()) start = ih->second;
      Lot l;
      l.shares = sh; l.basis = sh > 0 ? cost / sh : price[i]; l.start = start > 0 ? start : d;
      book[a].push_back(l);
    } else if (action[i] == 0) {
      double need = shares[i];
      const double px = price[i];
      std::vector<Lot>& lots = book[a];
      const std::vector<int>& bd = buy_days[a];
      const auto in_window = std::upper_bound(bd.begin(), bd.end(), d + 30) - std::lower_bound(bd.begin(), bd.end(), d - 30);
      const auto same_day = std::upper_bound(bd.begin(), bd.end(), d) - std::lower_bound(bd.begin(), bd.end(), d);
      const bool rebought = in_window > same_day;
      lots_to_sell(lots, need, lot_method, d, px, pick);
      for (int k : pick) {
        if (need <= kEpsShares) break;
        Lot& lot = lots[k];
        if (lot.shares <= kEpsShares) continue;
        const double sold = std::min(need, lot.shares);
        const double gain = sold * px - sold * lot.basis;
        if (gain < 0 && rebought) { if (y) y[4] += std::fabs(gain); }
        else if (y) {
          const bool lt = d - lot.start > 365;
          if (gain > 0) y[lt ? 2 : 0] += gain;
          else y[lt ? 3 : 1] += std::fabs(gain);
        }
        lot.shares -= sold;
        need -= sold;
      }
      lots.erase(std::remove_if(lots.begin(), lots.end(), [](const Lot& l) { return !(l.shares > kEpsShares); }), lots.end());
    }
  }
  return LETF_OK;
}
