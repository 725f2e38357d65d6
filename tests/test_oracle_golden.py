"""The CPU oracle against the golden vectors generated from the LIVE reference
(oracle/make_golden.py -> tests/golden/*.json).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.common import GOLDEN_NAMES, assert_matches_reference, case_inputs, load_case


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_reproduces_reference(name):
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    out = orc.run(params, tape)
    # fp64 on both sides: differences only from pow()/tanh() library rounding (<= a few ulp)
    assert_matches_reference(out, doc, rtol=1e-12, dd_atol=1e-13)
    assert out.n_raised == 0


def test_voldrag_known_answer_arithmetic():
    """validation.py:45-57 on the same shocks: prod(1 + L r) - 1 for L = 3, 2, 1."""
    doc = load_case("voldrag_c1")
    ns, model, params, keep, tape = case_inputs(doc)
    out = orc.run(params, tape)
    kat = np.array(doc["meta"]["validation_py_annual_return"])       # [trial][L=3,2,1]
    got = out.final_wealth / params.initial_capital - 1.0            # strategies: B_UPRO(3x), S3(2x), S2(1x)
    np.testing.assert_allclose(got.T, kat, rtol=1e-12, atol=1e-14)


def test_golden_cover_every_strategy_and_generator():
    seen_types, markets = set(), set()
    for name in GOLDEN_NAMES:
        doc = load_case(name)
        ns, model, params, keep, tape = case_inputs(doc)
        markets.add(params.market_mode)
        for k in range(params.n_strategies):
            seen_types.add(params.strategies[k].type)
    assert markets == {0, 1, 2, 4}
    assert seen_types == set(range(18))


def test_trade_list_matches_counts():
    doc = load_case("inst_1y")
    ns, model, params, keep, tape = case_inputs(doc)
    out = orc.run(params, tape)
    sids = doc["case"]["strategies"]
    for sid in ("S7", "S8", "S19"):
        s = sids.index(sid)
        tl = orc.trade_list(params, tape, 1, s)
        assert len(tl) == out.journal_count[s, 1]
        assert abs(sum(t["dollar_amount"] for t in tl) - out.journal_volume[s, 1]) < 1e-6
