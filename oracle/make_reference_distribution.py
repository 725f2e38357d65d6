"""Generate tests/golden/ref_distribution_2y.npz: terminal statistics of the LIVE reference run with its OWN
NumPy random streams (no tape), for the KS / summary-quantile parity of the on-device Philox mode.

TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference).  The reference is
imported unmodified and `simulate_single_path_fixed` is called in-process (fork pool), because going
through `parallel_monte_carlo_fixed` would silently revert cfg flags inside loky workers.

usage: python oracle/make_reference_distribution.py [--pairs 1000] [--years 2] [--procs 8]
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
SIDS = ["S1", "S2", "S3", "S4", "S5", "S6", "S7", "S19", "B_QQQ", "B_TLT"]

from letf_monte_carlo_b200 import synthetic  # noqa: E402


MODE = "inst"          # "inst": institutional_v1 defaults; "boot": legacy_hybrid block bootstrap over synthetic_history()
_BOOT = {}


def _init():
    sys.path.insert(0, REF)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
    import warnings
    warnings.filterwarnings("ignore")
    from letf import config as ref_cfg
    for k, v in synthetic.EXTRA_BENCHMARKS.items():
        ref_cfg.STRATEGIES[k] = dict(v)
    if MODE == "anchor":      # random_start.py:182-191: start regime from a uniformly drawn historical row
        ref_cfg.RANDOM_START_METHOD = "historical_anchor"
        _BOOT["anchors"] = synthetic.synthetic_history(n_days=3000, seed=5)
    if MODE == "boot":
        import contextlib, io
        from letf.simulation.bootstrap import BlockBootstrapReturns
        ref_cfg.SIM_ENGINE_MODE = "legacy_hybrid"
        with contextlib.redirect_stdout(io.StringIO()):      # the reference's own sampler (bootstrap.py:27-47)
            _BOOT["sampler"] = BlockBootstrapReturns(synthetic.synthetic_history(), block_size=ref_cfg.BOOTSTRAP_BLOCK_SIZE)


def _one(args):
    sim_id, years, anti = args
    from letf.simulation import engine as ref_engine
    model = synthetic.fallback_regime_model("default")
    if MODE == "boot":
        model["bootstrap_sampler"] = _BOOT["sampler"]
    if MODE == "anchor":
        model["historical_df_for_anchors"] = _BOOT["anchors"]
    res = ref_engine.simulate_single_path_fixed((sim_id, years, model, {}, SIDS, anti))
    row = []
    for s in SIDS:
        r = res[s]
        assert "Error" not in r, r.get("Error")
        row.append((float(r["Final_Wealth"]), float(r["Max_DD"]), int(r["Num_Trades"])))
    md = res[SIDS[0]]["Metadata"]
    rp = np.asarray(res[SIDS[0]]["Regime_Path"])
    return row, int((rp == 1).sum()), int(md.get("stress_summary", {}).get("jump_days", 0)), \
        bool(md["layer_checks"].get("repaired_after_validation", False))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1000)
    ap.add_argument("--years", type=float, default=2)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--out", default=None)
    ap.add_argument("--mode", default="inst", choices=["inst", "boot", "anchor"])
    a = ap.parse_args()
    global MODE
    MODE = a.mode
    if a.out is None:
        a.out = os.path.join(ROOT, "tests", "golden", {"inst": "ref_distribution_2y.npz", "boot": "ref_distribution_boot_1y.npz",
                                                       "anchor": "ref_distribution_anchor_1y.npz"}[MODE])
    # mc_runner.py:145-153: (sim_id, False), (sim_id, True) for sim_id = 0, 2, 4, ...
    tasks = [(2 * k, a.years, anti) for k in range(a.pairs) for anti in (False, True)]
    if MODE == "boot":     # outside the joint-t generator the antithetic flag is never read: twins would be identical paths
        tasks = [(k, a.years, False) for k in range(2 * a.pairs)]
    t0 = time.time()
    with mp.get_context("fork").Pool(a.procs, initializer=_init) as pool:
        out = pool.map(_one, tasks, chunksize=4)
    n = len(out)
    fw = np.array([[o[0][s][0] for o in out] for s in range(len(SIDS))])
    dd = np.array([[o[0][s][1] for o in out] for s in range(len(SIDS))])
    nt = np.array([[o[0][s][2] for o in out] for s in range(len(SIDS))], dtype=np.int32)
    import pandas, scipy
    np.savez_compressed(a.out, strategies=np.array(SIDS), years=a.years, final_wealth=fw, max_dd=dd, num_trades=nt,
                        high_vol_days=np.array([o[1] for o in out], np.int32),
                        jump_days=np.array([o[2] for o in out], np.int32),
                        repaired=np.array([o[3] for o in out]),
                        versions=np.array([f"numpy {np.__version__}", f"pandas {pandas.__version__}", f"scipy {scipy.__version__}"]),
                        generator=np.array("oracle/make_reference_distribution.py: live reference, its own PCG64 streams, "
                                           + ("fallback regime model, default cfg (institutional_v1, antithetic, moment matching, offset start)"
                                              if MODE == "inst" else
                                              "institutional_v1 defaults with RANDOM_START_METHOD = 'historical_anchor' over "
                                              "synthetic_history(n_days=3000, seed=5)" if MODE == "anchor" else
                                              "legacy_hybrid block bootstrap (the reference's BlockBootstrapReturns over synthetic_history()), offset start")))
    print(f"{n} trials in {time.time() - t0:.0f}s -> {a.out}")


if __name__ == "__main__":
    main()
