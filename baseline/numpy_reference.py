"""Time the reference's own NumPy CPU path (the unmodified package under baseline/_ref/) on this box's host cores.

`parallel_monte_carlo_fixed` (letf/mc_runner.py:24-210) for the path workloads, `validate_zero_drift_vol_drag`
(letf/validation.py:18-87) for config 1 at its literal size.  The reference cannot switch its market layers off: every
workload runs its full institutional_v1 engine (all six ETF layers) with the workload's benchmark-type strategy ids that
exist in its own cfg.STRATEGIES (loky workers re-import letf.config, so ids injected in the parent are not visible there).
Runs in a subprocess so that numba / joblib state and the reference's prints stay out of the bench process.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")

# strategy ids the reference itself defines that stand in for each workload's strategy set (config.py:348-437)
REF_SIDS = {"c2": ["S1", "S3"], "c3": ["S2"], "c4": ["S1", "S2", "S4", "S5", "S6", "S19"], "c5": ["S3", "S2"],
            "sweep19": ["S%d" % i for i in range(1, 20)], "boot": ["S1", "S2", "S3"]}


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "letf"))


def _child(workload: str, years: float, trials: int, workers: int) -> None:
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_baseline")
    import contextlib
    import io
    import warnings
    warnings.filterwarnings("ignore")
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        from letf import config as cfg
        if workload == "c1":
            from letf import validation
            t0 = time.perf_counter()
            validation.validate_zero_drift_vol_drag()
            dt = time.perf_counter() - t0
            res = {"seconds": dt, "path_days": 10_000 * 252, "trials": 10_000, "workers": 1,
                   "entry": "letf.validation.validate_zero_drift_vol_drag (10 000 x 252, vectorised NumPy)"}
        else:
            from letf import mc_runner
            from letf_monte_carlo_b200 import synthetic
            model = synthetic.fallback_regime_model("default")
            sids = REF_SIDS.get(workload, ["S2"])
            cfg.N_WORKERS = workers
            cfg.NUM_SIMULATIONS = max(2, min(trials, 2 * workers))     # warm-up: worker spin-up, imports, numba caches
            mc_runner.parallel_monte_carlo_fixed(sids, 1, model, {}, None)
            cfg.NUM_SIMULATIONS = trials
            t0 = time.perf_counter()
            out = mc_runner.parallel_monte_carlo_fixed(sids, years, model, {}, None)
            dt = time.perf_counter() - t0
            n = len(out[sids[0]])
            res = {"seconds": dt, "path_days": n * int(years * 252), "trials": n, "workers": workers, "strategies": sids,
                   "entry": "letf.mc_runner.parallel_monte_carlo_fixed (institutional_v1, all defaults)"}
    print(json.dumps(res), flush=True)


def numpy_rate(workload: str, years: float, trials: int, workers: int, timeout_s: float = 600.0) -> dict:
    """path-days/s of the unmodified reference; {'unavailable': why} when baseline/_ref is absent or the run fails"""
    if not available():
        return {"unavailable": "baseline/_ref/letf not present (python baseline/install_reference.py)"}
    cmd = [sys.executable, os.path.abspath(__file__), workload, str(years), str(trials), str(workers)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"reference run exceeded {timeout_s:.0f} s"}
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"unavailable": "reference run failed: " + (r.stderr.strip().splitlines() or ["no output"])[-1][:200]}
    d = json.loads(lines[-1])
    d["value"] = d["path_days"] / d["seconds"]
    return d


if __name__ == "__main__":
    _child(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
