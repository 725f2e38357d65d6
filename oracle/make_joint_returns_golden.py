"""tests/golden/joint_returns_400d.npz: the LIVE reference's simulate_joint_returns_t (letf/calibration.py:480-590) along a
caller-supplied regime path, plain and antithetic, with and without moment matching, fed with pre-drawn variates through
the TapeGenerator of oracle/make_golden.py (chol(cov_t) @ x for the correlated normals: the documented deviation from
NumPy's SVD factor).  TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from letf_monte_carlo_b200 import synthetic  # noqa: E402
from letf_monte_carlo_b200.tape import HostTape  # noqa: E402
# (the reference is imported only when this script runs as the generator: tests import `inputs` on boxes without it)


def inputs(n_days=400, seed=21):
    rng = np.random.default_rng(seed)
    regime = np.zeros(n_days, dtype=int)
    regime[60:95] = 1; regime[200:203] = 1; regime[330:] = 1            # starts low-vol; a 3-day spell; ends high-vol
    x = rng.standard_normal((n_days, 3))
    x[250] *= 90.0                                                      # a shock large enough to hit the -0.95 / 4.0 clips
    joint = synthetic.fallback_joint_model()
    nu = np.where(regime == 0, joint["regimes"][0]["nu"], joint["regimes"][1]["nu"])
    chi2 = rng.chisquare(nu)
    return regime, x, chi2, joint


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference")
    from letf import calibration as ref_cal
    from letf import config as ref_cfg
    from oracle.make_golden import TapeGenerator
    regime, x, chi2, joint = inputs()
    n = len(regime)
    tape = HostTape(1, n, n, 1)
    tape.x[:, :, 0] = x
    tape.chi2[:, 0] = chi2
    out = {"regime": regime, "x": x, "chi2": chi2}
    for mm in (True, False):
        ref_cfg.USE_MOMENT_MATCHING = mm
        for anti in (False, True):
            g = TapeGenerator("B", tape, 0, regime_full=regime)
            r = ref_cal.simulate_joint_returns_t(n, regime, joint, g, antithetic=anti)
            out[f"ret_mm{int(mm)}_anti{int(anti)}"] = np.stack([r[a] for a in joint["assets"]], axis=1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "joint_returns_400d.npz"), **out)
    print({k: v.shape for k, v in out.items()}, float(np.abs(out["ret_mm1_anti0"]).max()))
