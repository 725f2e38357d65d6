"""Host tax engine for the sampled-trial tax path: what the reference's
`process_trades_with_wired_engine` (letf/integration.py:26-556) computes for ONE trade list.

It is host bookkeeping on <= 50 trade lists per taxable strategy (reporting.py:455-467), outside the GPU path; it lives here
so that `taxable_summary` has no run-time dependency on the reference package.  Stages, each citing what it restates:

  1. wash sales, look-back and look-forward (letf/tax/wash_sale.py:149-397, IRC 1091): which loss sales have a replacement buy
     within +-30 trading days, how much loss is disallowed, which replacement lot takes the basis and the holding period;
  2. lot bookkeeping with the taxpayer's lot-selection election (integration.py:124-226, letf/tax/lot_selection.py:5-180):
     realised short / long-term gains and losses per tax year (252 trading days);
  3. yearly netting with loss carry-forwards and the $3 000 deduction (letf/tax/engine.py:80-324, IRC 1222 / 1211(b) / 1212(b));
  4. the ordinary-income path of the taxpayer (integration.py:238-337, seeded career model) and the margin-interest deduction
     (:363-384, :449-452);
  5. incremental federal + state + NIIT tax of the year's gains on top of that income, with marginal brackets inflated 2.5 % a
     year and long-term gains stacked on ordinary income (letf/tax/marginal.py:9-232, bracket tables letf/tax/brackets.py).

Reference behaviours kept on purpose (each one changes the result): every sell's wash-sale cost basis is FIFO over ALL of the
asset's buys, earlier sells never deplete them (wash_sale.py:208-209); a replacement buy serves one loss sale only; a
holding-period adjustment to day 0 is ignored (:451-456); the basis adjustment is keyed by the replacement's buy DAY, so
two buys of one day both receive it; at sale time a loss is dropped entirely when ANY other-day buy of the asset lies within
+-30 days (integration.py:196-207); the long-term threshold compares TRADING days with 365 (:209).

Pinned to the live reference: its six hand-calculated netting cases (tax/engine.py:405-508), 650 device-replayable trade lists
with the totals the reference computed for them (tests/golden/tax/tax_sampling_2y.json) and a synthetic stress set over
every election, lot method, filing status and state (tests/golden/tax/engine_cases.json, oracle/make_tax_engine_golden.py).
"""
from __future__ import annotations

from bisect import bisect_left, bisect_right
from collections import defaultdict
from itertools import accumulate
from typing import Mapping, Sequence

import numpy as np

INF = float("inf")
DAYS_PER_YEAR = 252
EPS_SHARES = 0.001

# ---- bracket tables (letf/tax/brackets.py: 2024 schedules; (upper limit, rate)) -----------------------------------------------
FEDERAL = {"single": [(11600, 0.10), (47150, 0.12), (100525, 0.22), (191950, 0.24), (243725, 0.32), (609350, 0.35), (INF, 0.37)],
           "married": [(23200, 0.10), (94300, 0.12), (201050, 0.22), (383900, 0.24), (487450, 0.32), (731200, 0.35), (INF, 0.37)]}
LTCG = {"single": [(47025, 0.00), (518900, 0.15), (INF, 0.20)], "married": [(94050, 0.00), (583750, 0.15), (INF, 0.20)]}
STD_DEDUCTION = {"single": 14600, "married": 29200}
NIIT_THRESHOLD = {"single": 200000, "married": 250000}
NIIT_RATE = 0.038
_NONE = {"single": [(INF, 0.0)], "married": [(INF, 0.0)], "std_deduction": {"single": 0, "married": 0}}
STATE = {
    "CA": {"single": [(10412, 0.01), (24684, 0.02), (38959, 0.04), (54081, 0.06), (68350, 0.08), (349137, 0.093), (418961, 0.103),
                      (698271, 0.113), (INF, 0.133)],
           "married": [(20824, 0.01), (49368, 0.02), (77918, 0.04), (108162, 0.06), (136700, 0.08), (698274, 0.093), (837922, 0.103),
                       (1396542, 0.113), (INF, 0.133)],
           "std_deduction": {"single": 5363, "married": 10726}},
    "NY": {"single": [(8500, 0.04), (11700, 0.045), (13900, 0.0525), (80650, 0.055), (215400, 0.06), (1077550, 0.0685),
                      (5000000, 0.0965), (25000000, 0.103), (INF, 0.109)],
           "married": [(17150, 0.04), (23600, 0.045), (27900, 0.0525), (161550, 0.055), (323200, 0.06), (2155350, 0.0685),
                       (5000000, 0.0965), (25000000, 0.103), (INF, 0.109)],
           "std_deduction": {"single": 8000, "married": 16050}},
    "TX": _NONE, "FL": _NONE, "NV": _NONE,
    "WA": {"single": [(250000, 0.0), (INF, 0.07)], "married": [(250000, 0.0), (INF, 0.07)],
           "std_deduction": {"single": 0, "married": 0}, "cap_gains_only": True},
    "IL": {"single": [(INF, 0.0495)], "married": [(INF, 0.0495)], "std_deduction": {"single": 2425, "married": 4850}},
    "MA": {"single": [(INF, 0.05)], "married": [(INF, 0.05)], "std_deduction": {"single": 0, "married": 0}},
    "NJ": {"single": [(20000, 0.014), (35000, 0.0175), (40000, 0.035), (75000, 0.05525), (500000, 0.0637), (1000000, 0.0897),
                      (INF, 0.1075)],
           "married": [(20000, 0.014), (50000, 0.0175), (70000, 0.0245), (80000, 0.035), (150000, 0.05525), (500000, 0.0637),
                       (1000000, 0.0897), (INF, 0.1075)],
           "std_deduction": {"single": 1000, "married": 2000}},
}

LOSS_STRATEGIES = ("offset_st_first", "offset_lt_first", "defer_maximum", "use_all_asap")     # tax/engine.py:21-26
LOT_METHODS = ("fifo", "lifo", "hifo", "lofo", "ltfo", "stfo", "mintax", "specific")          # tax/engine.py:45-52


# ---- 3. yearly netting (tax/engine.py:80-324) -------------------------------------------------------------------------------
def net_capital_gains(st_gains, st_losses, lt_gains, lt_losses, st_cf_in=0.0, lt_cf_in=0.0, strategy="offset_st_first") -> dict:
    """One tax year: net each class, cross-net the current year, then apply the carry-forwards in the order the election asks
    for, deduct up to $3 000 of a remaining net loss (short-term first) and carry the rest."""
    st, lt = st_gains - st_losses, lt_gains - lt_losses
    if st > 0 and lt < 0:                       # a gain in one class absorbs the other class's current-year loss
        x = min(st, -lt); st -= x; lt += x
    elif lt > 0 and st < 0:
        x = min(lt, -st); lt -= x; st += x
    cf = {"st": st_cf_in, "lt": lt_cf_in}
    net = {"st": st, "lt": lt}

    def use(src, dst):
        if cf[src] > 0 and net[dst] > 0:
            x = min(cf[src], net[dst]); net[dst] -= x; cf[src] -= x

    if strategy == "offset_lt_first":
        order = (("lt", "lt"), ("st", "st"), ("lt", "st"), ("st", "lt"))
    elif strategy == "defer_maximum":           # own class only, no cross-application
        order = (("st", "st"), ("lt", "lt"))
    else:                                       # offset_st_first and use_all_asap share one order
        order = (("st", "st"), ("lt", "lt"), ("st", "lt"), ("lt", "st"))
    for src, dst in order:
        use(src, dst)
    total = net["st"] + net["lt"]
    deduction = min(3000, abs(total)) if total < 0 else 0
    st_cf, lt_cf = cf["st"], cf["lt"]
    if net["st"] < 0:
        st_cf += max(0, abs(net["st"]) - deduction)
    if net["lt"] < 0:
        left = deduction - min(deduction, abs(min(0, net["st"])))        # what the short-term loss did not use of the $3 000
        lt_cf += max(0, abs(net["lt"]) - left)
    return {"taxable_st": max(0, net["st"]), "taxable_lt": max(0, net["lt"]), "st_loss_cf_out": st_cf, "lt_loss_cf_out": lt_cf,
            "capital_loss_deduction": deduction}


# ---- 5. marginal tax (tax/marginal.py) --------------------------------------------------------------------------------------
def marginal_tax(income, brackets, deduction=0.0) -> float:
    if income <= 0:
        return 0
    taxable = max(0, income - deduction)
    tax, lower = 0, 0
    for upper, rate in brackets:
        if taxable <= lower:
            break
        tax += (min(taxable, upper) - lower) * rate
        lower = upper
        if taxable <= upper:
            break
    return tax


def stacked_ltcg_tax(gain, ordinary_after_deduction, brackets) -> float:
    """long-term gains fill the preferential brackets starting where ordinary income ends (marginal.py:43-107)"""
    if gain <= 0:
        return 0.0
    pos, left, tax = max(0, ordinary_after_deduction), gain, 0.0
    for ceiling, rate in brackets:
        if left <= 0:
            break
        if pos >= ceiling:
            continue
        part = min(left, ceiling - pos)
        tax += part * rate
        left -= part
        pos += part
    return tax


def comprehensive_tax(taxable_st, taxable_lt, capital_loss_deduction, ordinary_income=0, filing_status="single",
                      bracket_multiplier=1.0, state_code="CA", include_state=True, include_niit=True) -> dict:
    """federal ordinary + short-term (as ordinary income) + stacked long-term + state + NIIT (marginal.py:110-232)"""
    fs = filing_status
    fed = FEDERAL.get(fs, FEDERAL["single"])
    std = STD_DEDUCTION.get(fs, STD_DEDUCTION["single"])
    ltcg = LTCG.get(fs, LTCG["single"])
    sd = STATE.get(state_code, STATE["CA"])
    sb = sd.get(fs, sd["single"])
    sstd = sd["std_deduction"].get(fs, sd["std_deduction"]["single"])
    niit_thr = NIIT_THRESHOLD.get(fs, NIIT_THRESHOLD["single"])
    if bracket_multiplier != 1.0:
        m = bracket_multiplier
        fed = [(t * m, r) for t, r in fed]
        ltcg = [(t * m, r) for t, r in ltcg]
        sb = [(t * m, r) for t, r in sb]
        std, niit_thr, sstd = std * m, niit_thr * m, sstd * m
    adj = max(0, ordinary_income - capital_loss_deduction)
    f_ord = marginal_tax(adj, fed, std)
    ord_total = adj + taxable_st
    f_st = marginal_tax(ord_total, fed, std) - f_ord
    total_income = ord_total + taxable_lt
    f_lt = stacked_ltcg_tax(taxable_lt, max(0, ord_total - std), ltcg) if taxable_lt > 0 else 0
    f_total = f_ord + f_st + f_lt
    s_tax = 0
    if include_state:
        s_tax = marginal_tax(taxable_st + taxable_lt if sd.get("cap_gains_only", False) else total_income, sb, sstd)
    n_tax = 0
    if include_niit and total_income > niit_thr:
        n_tax = min(taxable_st + taxable_lt, total_income - niit_thr) * NIIT_RATE
    return {"federal_ordinary": f_ord, "federal_st": f_st, "federal_ltcg": f_lt, "federal_total": f_total, "state_tax": s_tax,
            "niit_tax": n_tax, "total_tax": f_total + s_tax + n_tax}


# ---- 1. wash sales (tax/wash_sale.py:149-397) -------------------------------------------------------------------------------
def wash_sale_adjustments(trades: Sequence[Mapping]):
    """-> (basis adjustment, holding-period start) per (asset, replacement buy day)"""
    by_asset = defaultdict(list)
    for t in trades:
        sh = t.get("shares", t["dollar_amount"] / t["price"])
        by_asset[t["asset"]].append((t["day_index"], t["action"].upper(), sh, t["price"]))
    basis = defaultdict(float)
    hold = {}
    for asset, rows in by_asset.items():
        rows = sorted(rows, key=lambda r: r[0])                         # stable: same-day trades keep their order
        buys = [r for r in rows if r[1] == "BUY"]
        bdays = [b[0] for b in buys]                                    # ascending (rows are sorted by day)
        # FIFO cost of a sale over ALL buys of the asset (the reference never depletes them between sales): running sums
        # over the buys that hold shares, in the order the reference's loop visits them.  A strategy that trades daily for
        # 30 years has thousands of buys per asset; walking them per sale is quadratic (seconds per trade list).
        live = [(bd, bsh, bpx) for bd, _, bsh, bpx in buys if bsh > EPS_SHARES]
        cum_sh = list(accumulate(b[1] for b in live))
        cum_cost = list(accumulate(b[1] * b[2] for b in live))
        used = set()
        for day, action, shares, price in rows:
            if action != "SELL":
                continue
            cost, first_day = 0, None
            if shares > EPS_SHARES and live:
                first_day = live[0][0]
                m = bisect_left(cum_sh, shares)                         # first buy that covers what is left
                if m >= len(live):
                    cost = cum_cost[-1]
                else:
                    prev_sh = cum_sh[m - 1] if m else 0.0
                    cost = cum_cost[m - 1] if m else 0
                    left = shares - prev_sh
                    if left > EPS_SHARES:
                        cost += min(left, live[m][1]) * live[m][2]
            loss = -(shares * price - cost)
            if loss <= 0:
                continue
            # replacement buy: the nearest one within 30 days (not the same day, not used yet); ties: the earlier buy
            lo, hi = bisect_left(bdays, day - 30), bisect_right(bdays, day + 30)
            cand = [(abs(bdays[i] - day), i) for i in range(lo, hi) if bdays[i] != day and i not in used]
            if not cand:
                continue
            _, i = min(cand, key=lambda c: c[0])
            rday, _, rsh, _ = buys[i]
            disallowed = loss if rsh >= shares else loss * (rsh / shares)
            basis[(asset, rday)] += disallowed
            hold[(asset, rday)] = first_day if first_day is not None else day
            used.add(i)
    return basis, hold


# ---- 2. lot selection (tax/lot_selection.py) --------------------------------------------------------------------------------
def lots_to_sell(lots, shares_needed, method, sale_day, sale_price) -> list:
    live = [i for i, l in enumerate(lots) if l["shares"] > EPS_SHARES]
    age = lambda i: sale_day - lots[i]["start"]
    if method == "lifo":
        order = live[::-1]
    elif method in ("hifo", "lofo"):
        order = sorted(live, key=lambda i: lots[i]["basis"], reverse=method == "hifo")
    elif method == "ltfo":
        order = sorted([i for i in live if age(i) > 365], key=age, reverse=True) + \
                sorted([i for i in live if age(i) <= 365], key=age, reverse=True)
    elif method == "stfo":
        order = sorted([i for i in live if age(i) <= 365], key=age) + sorted([i for i in live if age(i) > 365], key=age)
    elif method in ("mintax", "specific"):
        def impact(i):
            g = sale_price - lots[i]["basis"]
            return g * ((0.20 if age(i) > 365 else 0.37) if g >= 0 else 0.37)
        order = sorted(live, key=impact)
    else:
        order = live
    out, left = [], shares_needed
    for i in order:
        if left <= EPS_SHARES:
            break
        out.append(i)
        left -= lots[i]["shares"]
    return out


_YEAR_KEYS = ("st_gains", "st_losses", "lt_gains", "lt_losses", "wash_sale_disallowed")


def realised_by_year_native(trades: Sequence[Mapping], lot_method="fifo"):
    """`realised_by_year` through the library's host routine (csrc/letf_tax.cu: the same bookkeeping, operation for operation);
    None when the journal is not plain BUY / SELL entries or the library is not built (the Python statement below is then used)."""
    from .trade import TradeList
    n = len(trades)
    columns = isinstance(trades, TradeList)            # a replayed journal: its columns as they are
    if n == 0 or (not columns and any(t["action"] not in ("BUY", "SELL") for t in trades)):
        return None
    try:
        from .abi import load_engine_library
        lib = load_engine_library()
    except Exception:
        return None
    if columns:
        day = np.ascontiguousarray(trades.day, np.int32)
        asset = np.ascontiguousarray(trades.asset, np.int32)
        action = np.ascontiguousarray((np.asarray(trades.action) > 0).astype(np.int32))
        price = np.ascontiguousarray(trades.price, np.float64)
        shares = np.ascontiguousarray(trades.shares, np.float64)
    else:
        ids = {}
        day = np.fromiter((t["day_index"] for t in trades), np.int32, n)
        asset = np.fromiter((ids.setdefault(t["asset"], len(ids)) for t in trades), np.int32, n)
        action = np.fromiter((1 if t["action"] == "BUY" else 0 for t in trades), np.int32, n)
        price = np.fromiter((t["price"] for t in trades), np.float64, n)
        shares = np.fromiter((t["shares"] if "shares" in t else t["dollar_amount"] / t["price"] for t in trades), np.float64, n)
    if day.min() < 0 or asset.min() < 0:
        return None
    n_years = int(day.max()) // DAYS_PER_YEAR + 1
    out = np.zeros((n_years, 5))
    rc = lib.letf_tax_realised_by_year(day.ctypes.data, asset.ctypes.data, action.ctypes.data, shares.ctypes.data,
                                       price.ctypes.data, n, LOT_METHODS.index(lot_method), n_years, out.ctypes.data)
    if rc != 0:
        return None
    years = defaultdict(lambda: {k: 0 for k in _YEAR_KEYS})
    for y in np.unique(day // DAYS_PER_YEAR).tolist():          # the years the journal touches, like the Python statement
        years[int(y)] = dict(zip(_YEAR_KEYS, out[int(y)].tolist()))
    return years


def realised_by_year(trades: Sequence[Mapping], lot_method="fifo", native: bool | None = None) -> dict:
    """{year: {st_gains, st_losses, lt_gains, lt_losses, wash_sale_disallowed}} (integration.py:113-226).
    native: None = the library's host routine when it applies (journals of a few hundred trades and more), True / False force it"""
    if native or (native is None and len(trades) >= 64):
        r = realised_by_year_native(trades, lot_method)
        if r is not None:
            return r
        if native:
            raise RuntimeError("native tax bookkeeping unavailable for this journal")
    basis_adj, hold = wash_sale_adjustments(trades)
    buy_days = defaultdict(list)
    for t in trades:
        if t["action"] == "BUY":
            buy_days[t["asset"]].append(t["day_index"])
    for d in buy_days.values():
        d.sort()
    years = defaultdict(lambda: {"st_gains": 0, "st_losses": 0, "lt_gains": 0, "lt_losses": 0, "wash_sale_disallowed": 0})
    book = defaultdict(list)
    for t in trades:
        asset, day = t["asset"], t["day_index"]
        y = years[day // DAYS_PER_YEAR]
        if t["action"] == "BUY":
            sh = t.get("shares", t["dollar_amount"] / t["price"])
            cost = sh * t["price"] + basis_adj.get((asset, day), 0.0)
            start = hold.get((asset, day), 0)
            book[asset].append({"shares": sh, "basis": cost / sh if sh > 0 else t["price"], "start": start if start > 0 else day})
        elif t["action"] == "SELL":
            need = t.get("shares", t["dollar_amount"] / t["price"])
            px, lots = t["price"], book[asset]
            # any buy of the asset within 30 days that is not on the sale's own day (bisection on the sorted buy days)
            bd = buy_days[asset]
            rebought = (bisect_right(bd, day + 30) - bisect_left(bd, day - 30)) > (bisect_right(bd, day) - bisect_left(bd, day))
            for i in lots_to_sell(lots, need, lot_method, day, px):
                if need <= EPS_SHARES:
                    break
                lot = lots[i]
                if lot["shares"] <= EPS_SHARES:
                    continue
                sold = min(need, lot["shares"])
                gain = sold * px - sold * lot["basis"]
                if gain < 0 and rebought:
                    y["wash_sale_disallowed"] += abs(gain)
                else:
                    cls = "lt" if day - lot["start"] > 365 else "st"
                    y[cls + ("_gains" if gain > 0 else "_losses")] += abs(gain) if gain <= 0 else gain
                lot["shares"] -= sold
                need -= sold
            book[asset] = [l for l in lots if l["shares"] > EPS_SHARES]
    return years


# ---- 4. ordinary income path (integration.py:238-337) -----------------------------------------------------------------------
def income_path(years: int, base_income: float, years_until_retirement=None, retirement_income=None):
    """seeded career model: 4 % base growth, N(0, 8 %) variation, setbacks / job changes / promotions, retirement drop.
    -> (income per year, index 0 = today; peak working income)"""
    rng = np.random.default_rng(42)
    path, recovery, peak = [base_income], 0, base_income
    for y in range(1, years + 1):
        cur = path[-1]
        if years_until_retirement is not None and y > years_until_retirement:
            if y == years_until_retirement + 1:
                new = peak * 0.50 if retirement_income is None else retirement_income
            else:
                new = cur * 1.02
        else:
            var = rng.normal(0, 0.08)
            event = 0
            if rng.random() < 0.03 and recovery == 0:
                event = rng.uniform(-0.20, -0.10)
                recovery = 2
            elif rng.random() < 0.10:
                event = rng.uniform(0.10, 0.30)
            elif rng.random() < 0.25:
                event = rng.uniform(0.15, 0.25)
            rec = 0
            if recovery > 0:
                rec = 0.08
                recovery -= 1
            growth = np.clip(0.04 + var + event + rec, -0.50, 1.00)
            new = cur * (1 + growth)
            peak = max(peak, new)
        path.append(new)
    return path, peak


# ---- the whole path (integration.py:26-556) ---------------------------------------------------------------------------------
def process_trades(trades: Sequence[Mapping], time_horizon_years, initial_capital: float, tax_config: Mapping | None = None,
                   lot_method: str = "fifo", loss_strategy: str = "offset_st_first") -> dict:
    """Tax analysis of one trade list: {'total_tax', 'yearly_taxes', 'total_st_gains', 'total_lt_gains', 'final_cf',
    'income_trajectory'} with the reference's meanings."""
    cfg = tax_config or {"filing_status": "single", "state": "CA", "ordinary_income": 150000, "career_stage": "mid"}
    H = int(time_horizon_years)
    if lot_method not in LOT_METHODS or loss_strategy not in LOSS_STRATEGIES:
        raise ValueError("unknown lot_method / loss_strategy")
    activity = realised_by_year(trades, lot_method)
    until, r_income = cfg.get("years_until_retirement", None), cfg.get("retirement_income", None)
    incomes, peak = income_path(H, cfg.get("ordinary_income", 150000), until, r_income)
    per_year = len(trades) / H if H > 0 else 0
    leverage = 1.2 if per_year < 50 else (1.5 if per_year < 150 else 1.8)
    base_margin = initial_capital * (leverage - 1) * 0.06
    fs, state = cfg.get("filing_status", "single").lower(), cfg.get("state", "CA")
    total, st_cf, lt_cf, rows = 0, 0, 0, []
    for year in range(H):
        a = activity[year]
        n = net_capital_gains(a["st_gains"], a["st_losses"], a["lt_gains"], a["lt_losses"], st_cf, lt_cf, loss_strategy)
        mult = (1 + 0.025) ** year
        margin = base_margin * (1 + (year * 0.15))
        st_after = max(0, n["taxable_st"] - margin)
        lt_after = max(0, n["taxable_lt"] - max(0, margin - n["taxable_st"]))
        kw = dict(capital_loss_deduction=n["capital_loss_deduction"], ordinary_income=incomes[year + 1], filing_status=fs,
                  bracket_multiplier=mult, state_code=state)
        base = comprehensive_tax(0, 0, **kw)
        full = comprehensive_tax(st_after, lt_after, **kw)
        tax = max(0, full["total_tax"] - base["total_tax"])
        total += tax
        st_cf, lt_cf = n["st_loss_cf_out"], n["lt_loss_cf_out"]
        gains = n["taxable_st"] + n["taxable_lt"]
        rows.append({"year": year, "taxable_st": n["taxable_st"], "taxable_lt": n["taxable_lt"], "tax": tax, "st_cf": st_cf, "lt_cf": lt_cf,
                     "capital_loss_deduction": n["capital_loss_deduction"], "federal_total": full["federal_total"] - base["federal_total"],
                     "state_tax": full["state_tax"] - base["state_tax"], "niit_tax": full["niit_tax"] - base["niit_tax"],
                     "effective_rate": tax / gains if gains > 0 else 0,
                     "wash_sale_disallowed": a["wash_sale_disallowed"]})
    return {"total_tax": total, "yearly_taxes": rows, "total_st_gains": sum(r["taxable_st"] for r in rows),
            "total_lt_gains": sum(r["taxable_lt"] for r in rows), "final_cf": {"st": st_cf, "lt": lt_cf},
            "income_trajectory": {"years": list(range(len(incomes))), "income": incomes, "peak_income": peak, "retirement_year": until}}
