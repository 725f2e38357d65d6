"""Golden for the sampled-trial tax path: the LIVE reference's `create_summary_statistics`
(letf/reporting.py:299-634) run on a result set of the reference's own shape, its printed taxable table parsed.

TEST INFRASTRUCTURE, authoring container only (needs /root/reference).  The per-trial results come from the CPU
oracle on a seeded tape (pinned to the reference by tests/golden/*.json); Trade_Lists from `oracle.trade_list`.
Stored: the case descriptor, the reference tax engine's total_tax for every (strategy, trial) it was asked about
(so that the test can replay the engine as a lookup - the reference is not importable where the tests run), and
the numbers the reference printed (dollars / percent at print precision).

usage: python oracle/make_tax_golden.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import make_golden as mg  # noqa: E402
from oracle import oracle as orc      # noqa: E402

CASE = dict(name="tax_sampling_2y", variant="default", years=2, n_tapes=36, seed=321, cfg={},
            strategies=["S2"] + ["S%d" % i for i in range(7, 20)])
TAX_CONFIG = {"filing_status": "single", "state": "CA", "state_name": "California", "ordinary_income": 150000,
              "career_stage": "mid"}


def build_results(case):
    """{sid: [reference-style dict per trial]} from the oracle."""
    ns, model, params, keep, tape = mg.case_inputs(case)
    out = orc.run(params, tape)
    n = out.final_wealth.shape[1]
    res = {}
    for s, sid in enumerate(case["strategies"]):
        rows = []
        for i in range(n):
            d = {"Final_Wealth": float(out.final_wealth[s, i]), "Max_DD": float(out.max_dd[s, i]),
                 "Num_Trades": int(out.num_trades[s, i]), "Trades_Per_Year": float(out.num_trades[s, i]) / case["years"]}
            if sid != "S2":
                d["Trade_List"] = orc.trade_list(params, tape, i, s)
            rows.append(d)
        res[sid] = rows
    return res


def main():
    sys.path.insert(0, REF)
    from letf import reporting, integration
    res = build_results(CASE)
    calls = {}
    orig = integration.process_trades_with_wired_engine

    def spy(trades, time_horizon_years, elections, initial_capital, debug=False, strategy_id="UNKNOWN", tax_config=None):
        r = orig(trades=trades, time_horizon_years=time_horizon_years, elections=elections,
                 initial_capital=initial_capital, debug=debug, strategy_id=strategy_id, tax_config=tax_config)
        key = f"{strategy_id}|{len(trades)}|{sum(t['dollar_amount'] for t in trades):.6f}"
        calls[key] = float(r["total_tax"])
        return r
    integration.process_trades_with_wired_engine = spy
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        reporting.create_summary_statistics(res, CASE["years"], tax_config=TAX_CONFIG)
    integration.process_trades_with_wired_engine = orig
    text = buf.getvalue()
    # main taxable table rows: rank id name $pre $post cagr% win% | maxdd% trd/y
    rows = {}
    pat = re.compile(r"^\s*(\d+)\s+(S\d+)\s+(.+?)\s+\$\s*([\d,\-]+)\s+\$\s*([\d,\-]+)\s+([\-\d.]+|nan)%\s+([\d.]+)%\s+\|\s+([\d.]+)%\s+([\d.]+)\s*$")
    for line in text.splitlines():
        m = pat.match(line)
        if m and m.group(2) in CASE["strategies"][1:]:
            rows[m.group(2)] = {"rank": int(m.group(1)), "pre_p60": float(m.group(4).replace(",", "")),
                                "post_p60": float(m.group(5).replace(",", "")),
                                "post_cagr_pct": float(m.group(6)), "win_pct": float(m.group(7)),
                                "max_dd_pct": float(m.group(8)), "trades_per_year": float(m.group(9))}
    assert len(rows) == 13, (len(rows), text[-3000:])
    import numpy, pandas, scipy
    doc = {"case": CASE, "tax_config": TAX_CONFIG, "tax_calls": calls, "printed": rows,
           "versions": {"numpy": numpy.__version__, "pandas": pandas.__version__, "scipy": scipy.__version__},
           "generator": "oracle/make_tax_golden.py (live reference reporting.create_summary_statistics + tax engine)"}
    with open(os.path.join(ROOT, "tests", "golden", "tax", "tax_sampling_2y.json"), "w") as f:
        json.dump(doc, f, indent=0)
    print("rows", len(rows), "tax calls", len(calls))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden", "tax"), exist_ok=True)
    main()
