"""Sampled-trial tax path against the LIVE reference's own report.

tests/golden/tax/tax_sampling_2y.json holds what `letf.reporting.create_summary_statistics` printed for a result
set built from the oracle, plus the reference tax engine's answer for every trade list it was handed
(oracle/make_tax_golden.py).  Here the same result set is rebuilt, the tax engine is replayed as a lookup, and
`taxable_summary` must select the same trials, interpolate the same taxes and land on the printed numbers."""
import json
import os

import numpy as np
import pytest

from letf_monte_carlo_b200.tax_sampling import interpolated_tax, tax_sample_indices, taxable_summary
from oracle import make_tax_golden as mtg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tax", "tax_sampling_2y.json")


@pytest.fixture(scope="module")
def doc():
    with open(GOLD) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def results(doc):
    return mtg.build_results(doc["case"])


def _lookup(doc, misses):
    calls = doc["tax_calls"]

    def fn(trades, years, capital, strategy_id, tax_config):
        key = f"{strategy_id}|{len(trades)}|{sum(t['dollar_amount'] for t in trades):.6f}"
        if key not in calls:
            misses.append(key)
            return 0.0
        return calls[key]
    return fn


def test_taxable_table_matches_the_reference_report(doc, results):
    misses = []
    rows = taxable_summary(results, doc["case"]["years"], tax_fn=_lookup(doc, misses), tax_config=doc["tax_config"])
    assert not misses                                   # exactly the trials the reference sampled, same trade lists
    printed = doc["printed"]
    assert [r["id"] for r in rows] == sorted(printed, key=lambda s: printed[s]["rank"])      # same ranking
    for r in rows:
        p = printed[r["id"]]
        assert abs(r["pcts"]["p60"]["pre_wealth"] - p["pre_p60"]) <= 0.5
        assert abs(r["pcts"]["p60"]["post_wealth"] - p["post_p60"]) <= 0.5
        assert abs(r["pcts"]["p60"]["post_cagr"] * 100 - p["post_cagr_pct"]) <= 0.00501
        assert abs(r["win"] - p["win_pct"]) <= 0.00501
        assert abs(r["max_dd"] * 100 - p["max_dd_pct"]) <= 0.00501
        assert abs(r["trades"] - p["trades_per_year"]) <= 0.00501
        assert len(r["sampled_trials"]) == 50


def test_sampling_and_interpolation_rules():
    """reporting.py:455-506: all trials when n <= 50 (exact taxes, no interpolation), evenly spaced ranks otherwise."""
    rng = np.random.default_rng(0)
    w = rng.lognormal(9, 1, 40)
    assert tax_sample_indices(w) == list(range(40))
    w = rng.lognormal(9, 1, 1000)
    idx = tax_sample_indices(w)
    order = np.argsort(w)
    assert len(idx) == 50 and idx[0] == order[0] and idx[-1] == order[-1]
    assert idx == [int(order[p]) for p in np.linspace(0, 999, 50, dtype=int)]
    res = [{"Final_Wealth": float(x), "Trade_List": [{"dollar_amount": float(x)}]} for x in w]
    tax = lambda trades, *a: 0.1 * trades[0]["dollar_amount"]          # noqa: E731  (linear in wealth -> interpolation exact)
    all_tax, idx2, st = interpolated_tax(res, "S7", 2, tax, None)
    np.testing.assert_allclose(all_tax, 0.1 * w, rtol=1e-12)
    res[3]["Trade_List"] = []                                          # no trades -> zero tax for that sampled trial
    small = res[:10]
    all_tax, idx3, st = interpolated_tax(small, "S7", 2, tax, None)
    assert idx3 == list(range(10)) and all_tax[3] == 0.0 and all_tax[4] == pytest.approx(0.1 * w[4])


def test_array_backed_results_replay_only_the_sampled_trials(doc, results):
    """StrategyResults-style views (arrays + trade_list(i)): the device replay is asked for 50 trials, not n."""
    sid = "S7"
    asked = []

    class View:
        final_wealth = np.array([r["Final_Wealth"] for r in results[sid]])
        max_dd = np.array([r["Max_DD"] for r in results[sid]])
        num_trades = np.array([r["Num_Trades"] for r in results[sid]])

        def __len__(self):
            return len(self.final_wealth)

        def trade_list(self, i):
            asked.append(i)
            return results[sid][i]["Trade_List"]

    misses = []
    rows_v = taxable_summary({sid: View(), "S2": results["S2"]}, doc["case"]["years"], tax_fn=_lookup(doc, misses),
                             tax_config=doc["tax_config"], taxable_ids=[sid])
    rows_d = taxable_summary({sid: results[sid], "S2": results["S2"]}, doc["case"]["years"], tax_fn=_lookup(doc, misses),
                             tax_config=doc["tax_config"], taxable_ids=[sid])
    assert not misses and sorted(asked) == sorted(rows_v[0]["sampled_trials"]) and len(asked) == 50
    assert rows_v[0]["pcts"] == rows_d[0]["pcts"] and rows_v[0]["win"] == rows_d[0]["win"]

    class BatchView(View):                 # the device path: ONE call for all sampled trials (letf_trade_lists_host)
        calls = []

        def trade_lists(self, idx):
            self.calls.append(list(idx))
            return [results[sid][i]["Trade_List"] for i in idx]

    bv = BatchView()
    rows_b = taxable_summary({sid: bv, "S2": results["S2"]}, doc["case"]["years"], tax_fn=_lookup(doc, misses),
                             tax_config=doc["tax_config"], taxable_ids=[sid])
    assert len(bv.calls) == 1 and bv.calls[0] == rows_b[0]["sampled_trials"] and rows_b[0]["pcts"] == rows_d[0]["pcts"]
