"""The host tax engine (letf_monte_carlo_b200/tax_engine.py) against the LIVE reference:
its six hand-calculated netting cases (letf/tax/engine.py:405-508), a synthetic stress set over every election / lot method /
filing status / state (tests/golden/tax/engine_cases.json, oracle/make_tax_engine_golden.py), and the 650 real trade lists of
the sampled-trial report fixture with the totals the reference computed for them (tests/golden/tax/tax_sampling_2y.json)."""
import json
import os

import numpy as np
import pytest

from letf_monte_carlo_b200 import tax_engine as te
from letf_monte_carlo_b200.tax_sampling import default_tax_fn, tax_sample_indices, taxable_summary

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tax")

# (st_gains, st_losses, lt_gains, lt_losses, st_cf_in, lt_cf_in, election) -> (taxable_st, taxable_lt, st_cf_out, lt_cf_out, deduction)
NETTING_GOLDENS = {        # letf/tax/engine.py:405-508, the reference's own hand-calculated cases
    "Basic Netting": ((50000, 10000, 20000, 5000, 0, 0, "offset_st_first"), (40000, 15000, 0, 0, 0)),
    "$3k Loss Deduction": ((5000, 20000, 0, 0, 0, 0, "offset_st_first"), (0, 0, 12000, 0, 3000)),
    "Cross-Offset Current Year": ((50000, 0, 0, 30000, 0, 0, "offset_st_first"), (20000, 0, 0, 0, 0)),
    "Carryforward Application Order": ((100000, 0, 0, 60000, 25000, 15000, "offset_st_first"), (0, 0, 0, 0, 0)),
    "Large Loss Year": ((10000, 500000, 5000, 200000, 0, 0, "offset_st_first"), (0, 0, 487000, 195000, 3000)),
    "Election Test: DEFER_TO_FUTURE": ((50000, 0, 30000, 0, 40000, 25000, "defer_maximum"), (10000, 5000, 0, 0, 0)),
}


@pytest.mark.parametrize("name", list(NETTING_GOLDENS))
def test_netting_golden_cases_of_the_reference(name):
    (sg, sl, lg, ll, scf, lcf, el), want = NETTING_GOLDENS[name]
    r = te.net_capital_gains(sg, sl, lg, ll, scf, lcf, el)
    got = (r["taxable_st"], r["taxable_lt"], r["st_loss_cf_out"], r["lt_loss_cf_out"], r["capital_loss_deduction"])
    assert all(abs(a - b) <= 0.01 for a, b in zip(got, want)), (got, want)


def test_ltcg_stacking_known_answer():
    """marginal.py:235-300: single filer, $40 000 ordinary income, $50 000 long-term gain -> $4 256.25"""
    tax = te.stacked_ltcg_tax(50000, 40000 - 14600, te.LTCG["single"])
    assert abs(tax - (50000 - (47025 - 25400)) * 0.15) < 1e-9


def test_stress_set_matches_live_reference():
    from oracle.make_tax_engine_golden import trade_list
    doc = json.load(open(os.path.join(GOLD, "engine_cases.json")))
    assert len(doc["cases"]) == 52          # 48 short journals + 4 at the benchmarked horizon (30 years, 3 000 - 6 000 trades)
    assert sum(d["case"]["years"] == 30 and d["total_tax"] > 0 for d in doc["cases"]) == 4
    seen = set()
    for d in doc["cases"]:
        c = d["case"]
        tl = trade_list(c["seed"], c["years"], c["n_assets"], c["n_trades"])
        r = te.process_trades(tl, c["years"], c["capital"], tax_config=c["config"], lot_method=c["lots"], loss_strategy=c["loss"])
        assert r["total_tax"] == pytest.approx(d["total_tax"], rel=1e-10, abs=1e-7), c
        np.testing.assert_allclose(r["income_trajectory"]["income"], d["income"], rtol=1e-13)
        assert len(r["yearly_taxes"]) == len(d["yearly"])
        for mine, ref in zip(r["yearly_taxes"], d["yearly"]):
            for k, v in ref.items():
                assert mine[k] == pytest.approx(v, rel=1e-10, abs=1e-7), (c, k)
        assert r["final_cf"]["st"] == pytest.approx(d["final_cf"]["st"], rel=1e-10, abs=1e-7)
        assert r["final_cf"]["lt"] == pytest.approx(d["final_cf"]["lt"], rel=1e-10, abs=1e-7)
        seen.add((c["loss"], c["lots"], c["config"]["state"], c["config"]["filing_status"]))
    assert len({s[0] for s in seen}) == 4 and len({s[1] for s in seen}) == 8 and len({s[2] for s in seen}) == 6


def test_real_trade_lists_match_the_totals_the_reference_computed():
    """every trade list the reference's report taxed in the sampled-trial fixture (13 strategies x 50 trials + the SPY leg):
    the engine here gives the reference's total to the cent - and the taxable table built with it lands on the printed report"""
    from oracle import make_tax_golden as mtg
    doc = json.load(open(os.path.join(GOLD, "tax_sampling_2y.json")))
    results = mtg.build_results(doc["case"])
    calls, checked = doc["tax_calls"], 0
    for sid, rows in results.items():
        if sid == "S2":
            continue
        fw = np.array([r["Final_Wealth"] for r in rows])
        for i in tax_sample_indices(fw):
            tl = rows[i].get("Trade_List") or []
            if not tl:
                continue
            key = f"{sid}_bulk|{len(tl)}|{sum(t['dollar_amount'] for t in tl):.6f}"
            if key not in calls:
                continue
            got = te.process_trades(tl, doc["case"]["years"], 10000.0, tax_config=doc["tax_config"])["total_tax"]
            assert got == pytest.approx(calls[key], rel=1e-9, abs=1e-6), key
            checked += 1
    assert checked >= 600
    rows = taxable_summary(results, doc["case"]["years"], tax_fn=None, tax_config=doc["tax_config"])     # default = this engine
    printed = doc["printed"]
    assert [r["id"] for r in rows] == sorted(printed, key=lambda s: printed[s]["rank"])
    for r in rows:
        p = printed[r["id"]]
        assert abs(r["pcts"]["p60"]["post_wealth"] - p["post_p60"]) <= 0.5
        assert abs(r["pcts"]["p60"]["post_cagr"] * 100 - p["post_cagr_pct"]) <= 0.00501


def test_default_tax_fn_is_the_in_tree_engine():
    fn = default_tax_fn()
    tl = [{"day_index": 10, "asset": "TQQQ", "action": "BUY", "shares": 100.0, "price": 50.0, "dollar_amount": 5000.0},
          {"day_index": 300, "asset": "TQQQ", "action": "SELL", "shares": 100.0, "price": 400.0, "dollar_amount": 40000.0}]
    assert fn(tl, 2, 10000.0, "S7_bulk", None) == te.process_trades(tl, 2, 10000.0)["total_tax"] > 0


def test_native_bookkeeping_equals_the_python_statement():
    """csrc/letf_tax.cu (host routine of the library) against `realised_by_year`'s Python statement: every lot method on the
    stress journals (incl. the four 30-year ones with thousands of trades), value for value"""
    from oracle.make_tax_engine_golden import trade_list
    doc = json.load(open(os.path.join(GOLD, "engine_cases.json")))
    checked = 0
    for d in doc["cases"]:
        c = d["case"]
        tl = trade_list(c["seed"], c["years"], c["n_assets"], c["n_trades"])
        for m in (te.LOT_METHODS if c["n_trades"] <= 400 else (c["lots"], "mintax")):
            a = te.realised_by_year(tl, m, native=True)
            b = te.realised_by_year(tl, m, native=False)
            assert sorted(a) == sorted(b)
            for y in b:
                for k, v in b[y].items():
                    assert a[y][k] == pytest.approx(v, rel=1e-13, abs=1e-9), (c, m, y, k)
            checked += 1
    assert checked >= 48 * 8
    # journals the routine does not take (lower-case actions are wash-sale buys but not lots in the reference): Python statement
    odd = [dict(t, action=t["action"].lower()) for t in trade_list(1003, 3, 2, 40)]
    assert te.realised_by_year_native(odd) is None
    assert te.process_trades(odd, 3, 10000.0)["total_tax"] >= 0.0


def test_trade_list_columns_read_like_the_reference_list_of_dicts():
    """`TradeList` (what the device replays return): list-of-dict behaviour for the reference's consumers, and the tax engine
    takes its columns directly - same totals as the same journal given as dicts"""
    from letf_monte_carlo_b200 import abi
    from letf_monte_carlo_b200.trade import TradeList
    from oracle.make_tax_engine_golden import trade_list
    dicts = [t for t in trade_list(2001, 30, 2, 5000)]
    for t in dicts:
        t.setdefault("shares", t["dollar_amount"] / t["price"])
        t["dollar_amount"] = t["shares"] * t["price"]
    names = abi.ASSET_NAMES
    tl = TradeList(np.array([t["day_index"] for t in dicts], np.int32), np.array([names.index(t["asset"]) for t in dicts], np.int32),
                   np.array([1 if t["action"] == "BUY" else -1 for t in dicts], np.int32), np.array([t["shares"] for t in dicts]),
                   np.array([t["price"] for t in dicts]), names)
    assert len(tl) == len(dicts) and bool(tl) and tl == dicts and not (tl == dicts[:-1])
    assert tl[0] == dicts[0] and tl[-1] == dicts[-1] and tl[3:6] == dicts[3:6] and list(tl)[17] == dicts[17]
    with pytest.raises(IndexError):
        tl[len(dicts)]
    a = te.process_trades(tl, 30, 10000.0)
    b = te.process_trades(dicts, 30, 10000.0)
    assert a["total_tax"] == pytest.approx(b["total_tax"], rel=1e-13) and a["total_tax"] > 0
    assert not TradeList(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), np.zeros(0), names)
