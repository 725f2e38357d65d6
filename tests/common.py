"""Shared helpers of the test-suite: golden-case loading and comparison rules."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from letf_monte_carlo_b200 import abi  # noqa: F401
from oracle import make_golden as mg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.json")))
GOLDEN_NAMES = [os.path.basename(f)[:-5] for f in GOLDEN_FILES]


def load_case(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        doc = json.load(f)
    return doc


def case_inputs(doc):
    """(config namespace, regime_model, params, keepalive, tape) of a golden case."""
    return mg.case_inputs(doc["case"])


def expected_arrays(doc):
    sids = doc["case"]["strategies"]
    exp = doc["expected"]
    fw = np.array([exp[s]["final_wealth"] for s in sids])
    dd = np.array([exp[s]["max_dd"] for s in sids])
    nt = np.array([exp[s]["num_trades"] for s in sids])
    jc = np.array([exp[s]["journal_count"] for s in sids])
    jv = np.array([exp[s]["journal_volume"] for s in sids])
    return fw, dd, nt, jc, jv


def assert_matches_reference(out, doc, rtol, dd_atol, exact_counts=True, journal=True):
    """out: HostOutputs-like.  Integer work (regime days, start offsets, trade and journal counts)
    must be identical; wealth within rtol (relative), drawdown within dd_atol (absolute)."""
    fw, dd, nt, jc, jv = expected_arrays(doc)
    meta = doc["meta"]
    np.testing.assert_allclose(out.final_wealth, fw, rtol=rtol, atol=0)
    np.testing.assert_allclose(out.max_dd, dd, rtol=0, atol=dd_atol)
    if exact_counts:
        np.testing.assert_array_equal(out.num_trades, nt)
        if journal and out.journal_count is not None:
            np.testing.assert_array_equal(out.journal_count, jc)
            np.testing.assert_allclose(out.journal_volume, jv, rtol=max(rtol, 1e-12), atol=1e-9)
    np.testing.assert_array_equal(out.start_offset, np.array(meta["start_offset"]))
    np.testing.assert_array_equal(out.regime_days.T, np.array(meta["regime_days"]))
    np.testing.assert_array_equal(out.jump_days, np.array(meta["jump_days"]))
    rep = (out.flags & abi.FLAG_REPAIRED) != 0
    np.testing.assert_array_equal(rep, np.array(meta["repaired"]))


def reference_distribution(name="ref_distribution_2y.npz"):
    """terminal statistics of the LIVE reference run with its own NumPy streams
    (oracle/make_reference_distribution.py -> tests/golden/ref_distribution_2y.npz, ref_distribution_boot_1y.npz)"""
    d = np.load(os.path.join(GOLDEN_DIR, name))
    return {k: d[k] for k in d.files}


def bootstrap_distribution_case():
    """(reference arrays, strategy ids, params, keepalive) of the legacy_hybrid block-bootstrap distribution fixture:
    the reference's own BlockBootstrapReturns over synthetic_history(), 8 000 independent trials, 1 year."""
    from letf_monte_carlo_b200 import synthetic
    from letf_monte_carlo_b200.params import build_params
    ref = reference_distribution("ref_distribution_boot_1y.npz")
    sids = [str(s) for s in ref["strategies"]]
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, SIM_ENGINE_MODE="legacy_hybrid",
                                    USE_ANTITHETIC_VARIATES=False)
    params, keep = build_params(sids, float(ref["years"]), synthetic.legacy_bootstrap_model(), ns)
    return ref, sids, params, keep


def anchor_distribution_case():
    """RANDOM_START_METHOD = 'historical_anchor' (random_start.py:182-191): the reference draws a uniform anchor row of the
    history and starts in its regime; here that is a Bernoulli start regime with the share of high-VIX valid rows."""
    from letf_monte_carlo_b200 import synthetic
    from letf_monte_carlo_b200.params import build_params
    ref = reference_distribution("ref_distribution_anchor_1y.npz")
    sids = [str(s) for s in ref["strategies"]]
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, RANDOM_START_METHOD="historical_anchor")
    model = synthetic.fallback_regime_model("default")
    model["historical_df_for_anchors"] = synthetic.synthetic_history(n_days=3000, seed=5)
    params, keep = build_params(sids, float(ref["years"]), model, ns)
    return ref, sids, params, keep


def assert_same_distribution(ref, out, sids, p_min=1e-3, rank_tol=4.0):
    """two-sample KS on terminal wealth and max drawdown per strategy, summary quantiles of wealth (rank agreement),
    mean trade counts within 3 %, regime / jump / repair statistics within sampling error"""
    from scipy import stats
    for k, s in enumerate(sids):
        a, b = ref["final_wealth"][k], out.final_wealth[k]
        assert stats.ks_2samp(a, b).pvalue > p_min, (s, "final wealth")
        assert stats.ks_2samp(ref["max_dd"][k], out.max_dd[k]).pvalue > p_min, (s, "max drawdown")
        for q in (10, 25, 40, 50, 60, 75, 90):       # reporting.py:361-366 percentiles
            # the reference's q-th percentile must sit at the q-th percentile of our sample, +- 4 points
            # (its own sampling error at n = 2 400 is ~ 1 point; wealth spans decades, so compare ranks)
            qa = np.percentile(a, q)
            rank = 100.0 * np.mean(b <= qa)
            assert abs(rank - q) < rank_tol, (s, q, qa, rank)
        ta, tb = ref["num_trades"][k].mean(), out.num_trades[k].mean()
        assert abs(tb - ta) <= 0.03 * max(ta, 1.0), (s, ta, tb)
    n_ref = len(ref["high_vol_days"])
    hv = out.regime_days[1]
    assert abs(hv.mean() - ref["high_vol_days"].mean()) < 5 * ref["high_vol_days"].std() / np.sqrt(n_ref / 2)
    assert abs(out.jump_days.mean() - ref["jump_days"].mean()) < 0.25
    rep = ((out.flags & abi.FLAG_REPAIRED) != 0).mean()
    assert abs(rep - ref["repaired"].mean()) < 0.06
