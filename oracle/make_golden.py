"""Generate tests/golden/*.json by running the LIVE reference on shock tapes.

TEST INFRASTRUCTURE.  Runs only in the authoring container (needs /root/reference,
read-only).  The reference code is imported unmodified; the only thing injected is
the random source: every `np.random.default_rng(seed)` the path constructs
(engine.py:43, 409, 436; random_start.py:157) and every `rng` it is handed
(calibration.py:480, 787; engine.py:157, 255) becomes a `TapeGenerator`, a subclass
of numpy's Generator that replays pre-drawn unit-scale variates from a HostTape.
One documented deviation: `multivariate_normal` returns
chol(cov) @ x_tape (numpy would factor with SVD), so that reference, oracle and
kernel share the Cholesky factor the north star prescribes.

The fixtures hold only the case descriptors (seed, sizes, config overrides) and the
reference's outputs; the tapes are re-drawn from the seed by
`letf_monte_carlo_b200.tape.make_tape` wherever the fixtures are used.

usage:  python oracle/make_golden.py [--only NAME] [--check]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from letf_monte_carlo_b200 import abi, synthetic  # noqa: E402
from letf_monte_carlo_b200.params import build_params  # noqa: E402
from letf_monte_carlo_b200.tape import make_tape  # noqa: E402

ALL19 = ["S%d" % i for i in range(1, 20)]
EXTRAS = list(synthetic.EXTRA_BENCHMARKS)

CASES = [
    dict(name="inst_1y", variant="default", years=1, n_tapes=6, seed=101, cfg={}, strategies=ALL19 + EXTRAS),
    dict(name="inst_2y_no_mm", variant="default", years=2, n_tapes=2, seed=102,
         cfg={"USE_MOMENT_MATCHING": False}, strategies=ALL19 + EXTRAS),
    dict(name="inst_1y_fixed_start", variant="default", years=1, n_tapes=3, seed=103,
         cfg={"USE_RANDOM_START": False}, strategies=ALL19),
    dict(name="no_stress_1y", variant="no_stress", years=1, n_tapes=2, seed=104, cfg={}, strategies=ALL19),
    dict(name="no_funding_1y", variant="no_funding", years=1, n_tapes=2, seed=105, cfg={}, strategies=ALL19),
    dict(name="durations_2y", variant="durations", years=2, n_tapes=2, seed=106, cfg={}, strategies=ALL19),
    dict(name="vix_te_1y", variant="vix_te", years=1, n_tapes=3, seed=107, cfg={}, strategies=ALL19 + EXTRAS),
    dict(name="param_t_1y", variant="no_joint", years=1, n_tapes=3, seed=108, cfg={}, strategies=ALL19 + EXTRAS),
    dict(name="legacy_param_t_1y", variant="default", years=1, n_tapes=2, seed=109,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid"}, strategies=ALL19),
    dict(name="jumpy_1y", variant="jumpy", years=1, n_tapes=3, seed=110, cfg={}, strategies=ALL19),
    dict(name="no_antithetic_1y", variant="default", years=1, n_tapes=3, seed=111,
         cfg={"USE_ANTITHETIC_VARIATES": False}, strategies=ALL19),
    dict(name="half_year", variant="default", years=0.5, n_tapes=2, seed=112, cfg={}, strategies=ALL19 + EXTRAS),
    dict(name="three_days", variant="default", years=3 / 252 + 1e-9, n_tapes=2, seed=113, cfg={}, strategies=ALL19),
    dict(name="one_day", variant="default", years=1 / 252 + 1e-9, n_tapes=1, seed=114, cfg={}, strategies=ALL19),
    dict(name="inst_10y", variant="default", years=10, n_tapes=2, seed=115, cfg={}, strategies=ALL19 + EXTRAS),
    dict(name="inst_30y", variant="default", years=30, n_tapes=1, seed=116, cfg={}, strategies=ALL19),
    dict(name="regime_only_start_1y", variant="default", years=1, n_tapes=4, seed=123,
         cfg={"RANDOM_START_METHOD": "regime_only", "START_REGIME_PROBABILITIES": {0: 0.4, 1: 0.6}}, strategies=ALL19),
    dict(name="boot_1y", variant="boot", years=1, n_tapes=3, seed=119,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid"}, strategies=ALL19 + EXTRAS),
    dict(name="boot_3y_nan_fixed_start", variant="boot_nan", years=3, n_tapes=2, seed=120,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid", "USE_RANDOM_START": False, "BOOTSTRAP_BLOCK_MEAN": 40,
              "BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME": {0: 0.9, 1: 0.2}}, strategies=ALL19),
    dict(name="boot_empty_tech_pool_1y", variant="boot_early", years=1, n_tapes=3, seed=125,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid"}, strategies=ALL19 + EXTRAS),
    dict(name="boot_no_highvol_blocks_1y", variant="boot_calm", years=1, n_tapes=3, seed=126,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid", "USE_RANDOM_START": False, "BOOTSTRAP_BLOCK_MEAN": 30}, strategies=ALL19),
    dict(name="boot_10y", variant="boot", years=10, n_tapes=1, seed=124,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid"}, strategies=ALL19 + EXTRAS),
    dict(name="boot_fullblocks_1y", variant="boot", years=1, n_tapes=2, seed=121,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid", "USE_RANDOM_START": False, "BOOTSTRAP_BLOCK_MIN": 168,
              "BOOTSTRAP_BLOCK_MAX": 168, "USE_ANTITHETIC_VARIATES": False}, strategies=ALL19),
    dict(name="boot_tinyblocks_halfyear", variant="boot_nan", years=0.5, n_tapes=2, seed=122,
         cfg={"SIM_ENGINE_MODE": "legacy_hybrid", "BOOTSTRAP_BLOCK_MIN": 1, "BOOTSTRAP_BLOCK_MAX": 5,
              "BOOTSTRAP_BLOCK_MEAN": 2, "BOOTSTRAP_WEIGHT": 1.0}, strategies=ALL19 + EXTRAS),
    dict(name="layers_c2_5y", variant="normal_c2", years=5, n_tapes=4, seed=117, cfg={},
         strategies=["S1", "S3", "B_UPRO"]),
    dict(name="layers_c2_30y", variant="normal_c2", years=30, n_tapes=2, seed=123, cfg={},
         strategies=["S1", "S3", "B_UPRO"]),
    dict(name="voldrag_c1", variant="normal_c1", years=1, n_tapes=8, seed=118,
         cfg={"LETF_CONFIGS": "pure_leverage"}, strategies=["B_UPRO", "S3", "S2"]),
]


def case_config(case):
    """our-config snapshot for a case (shared with the tests)."""
    over = dict(case["cfg"])
    if "START_REGIME_PROBABILITIES" in over:
        over["START_REGIME_PROBABILITIES"] = {int(k): float(v) for k, v in over["START_REGIME_PROBABILITIES"].items()}
    if "BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME" in over:          # JSON round trip turns the regime keys into strings
        over["BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME"] = {int(k): float(v) for k, v in over["BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME"].items()}
    if over.get("LETF_CONFIGS") == "pure_leverage":
        over["LETF_CONFIGS"] = synthetic.pure_leverage_assets()
        over["ASSETS"] = over["LETF_CONFIGS"]
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, **over)
    if case["variant"].startswith("normal_"):
        ns.USE_RANDOM_START = False
        ns.USE_ANTITHETIC_VARIATES = False
    return ns


def case_model(case):
    v = case["variant"]
    if v.startswith("normal_"):
        return synthetic.normal_regime_model(v.split("_", 1)[1])
    if v.startswith("boot"):
        if v == "boot_nan":
            hist = synthetic.synthetic_history(nan_every=97)
        elif v == "boot_early":        # ends before the real-QQQ cutoff: pool B is empty in both regimes
            hist = synthetic.synthetic_history(n_days=4000, seed=8, start="1980-01-02")
        elif v == "boot_calm":         # VIX never reaches 25: no high-vol blocks in either pool
            hist = synthetic.synthetic_history(n_days=9000, seed=10, start="1992-01-02")
            hist["VIX"] = hist["VIX"].clip(upper=22.0)
        else:
            hist = synthetic.synthetic_history()
        return synthetic.legacy_bootstrap_model(hist)
    return synthetic.fallback_regime_model(v)


def case_inputs(case):
    ns = case_config(case)
    model = case_model(case)
    params, keep = build_params(case["strategies"], case["years"], model, ns)
    tape = make_tape(params, case["n_tapes"], case["seed"])
    return ns, model, params, keep, tape


# --------------------------------------------------------------------------------------------
def _chol3(c):
    l00 = math.sqrt(c[0][0]); l10 = c[1][0] / l00; l20 = c[2][0] / l00
    l11 = math.sqrt(c[1][1] - l10 * l10); l21 = (c[2][1] - l20 * l10) / l11
    l22 = math.sqrt(c[2][2] - l20 * l20 - l21 * l21)
    return l00, l10, l11, l20, l21, l22


class TapeGenerator(np.random.Generator):
    """numpy Generator that replays a HostTape column; `role` selects the stream
    (SURVEY.md Appendix B): 'S' start, 'A' regime+stress, 'B' market, 'T<a>' tracking error."""

    def __init__(self, role, tape, col, regime_full=None):
        super().__init__(np.random.PCG64(0))
        self.role, self.tp, self.c = role, tape, col
        self.regime_full = regime_full
        self.k_spell = 0
        self.n_norm = 0
        self.n_rand = 0
        self.g_mvn = 0
        self.g_chi = 0
        self.g_vix = 1
        self.irx_r = 0
        self.i_te = 1

    # stream S -------------------------------------------------------------------------
    def integers(self, low, high=None, *a, **k):
        assert self.role == "S"
        v = int(self.tp.start_offset[self.c])
        assert low <= v < high
        return v

    def uniform(self, low=0.0, high=1.0, size=None):
        assert self.role == "S"
        return float(low)            # initial_vix: never read when VIX is generated

    # stream A -------------------------------------------------------------------------
    def _spell(self):
        v = int(self.tp.spells[self.k_spell, self.c]) if self.k_spell < self.tp.max_spells else self.tp.gen_days
        self.k_spell += 1
        return v

    def geometric(self, p, size=None):
        assert self.role == "A" and size is None
        return self._spell()

    def choice(self, a, size=None, replace=True, p=None, **k):
        if self.role == "S":                         # random_start.py:31 start regime ('regime_only')
            assert list(a) == [0, 1] and size is None
            return int(self.tp.start_regime[self.c])
        assert self.role == "A"
        if size is not None:                         # engine.py:509 jump signs
            assert list(a) == [-1.0, 1.0]
            return self.tp.sign[:size, self.c].copy()
        if p is not None:                            # engine.py:299 next regime
            return int(np.argmax(p))
        return self._spell()                         # engine.py:280 dwell-time sample

    def random(self, size=None, *a, **k):
        assert self.role == "A" and size is None
        self.n_rand += 1
        return float(self.tp.jump_u[self.n_rand, self.c])

    def standard_t(self, df, size=None):
        if self.role == "A":                         # calibration.py:828
            return float(self.tp.jump_t[self.n_rand, self.c])
        assert self.role.startswith("T")
        a = int(self.role[1:])
        v = float(self.tp.te_t[a, self.i_te, self.c])
        self.i_te += 1
        return np.float64(v) if size is None or size == () else np.full(size, v)

    def normal(self, loc=0.0, scale=1.0, size=None):
        if self.role == "A":                         # calibration.py:814-815
            day = 1 + self.n_norm // 2
            arr = self.tp.liq_n if self.n_norm % 2 == 0 else self.tp.cred_n
            self.n_norm += 1
            return loc + scale * float(arr[day, self.c])
        assert self.role == "B"
        if size is None:                             # engine.py:115
            v = float(self.tp.vix_n[self.g_vix, self.c])
            self.g_vix += 1
            return loc + scale * v
        reg = self.regime_full                       # engine.py:117-123
        while (reg == self.irx_r).sum() == 0:
            self.irx_r += 1
        m = reg == self.irx_r
        assert m.sum() == size
        self.irx_r += 1
        return loc + scale * self.tp.irx_n[m, self.c]

    # stream B -------------------------------------------------------------------------
    def multivariate_normal(self, mean, cov, *a, **k):
        assert self.role == "B"
        x0, x1, x2 = (float(v) for v in self.tp.x[self.g_mvn, :, self.c])
        self.g_mvn += 1
        l00, l10, l11, l20, l21, l22 = _chol3(cov)
        return np.array([mean[0] + l00 * x0, mean[1] + (l10 * x0 + l11 * x1),
                         mean[2] + (l20 * x0 + l21 * x1 + l22 * x2)])

    def chisquare(self, df, size=None):
        assert self.role == "B"
        v = float(self.tp.chi2[self.g_chi, self.c])
        self.g_chi += 1
        return v


class PatchedRng:
    """context manager: np.random.default_rng -> TapeGenerators bound to one tape column."""

    def __init__(self, tape, col, sim_id, regime_full, ref_engine, param_t=False, boot=None):
        self.boot = boot
        self.tape, self.col, self.sim_id, self.reg = tape, col, sim_id, regime_full
        self.eng = ref_engine
        self.calls_50000 = 0
        self.served = set()
        self.next_asset = None
        self.param_t = param_t

    def factory(self, seed=None):
        s = self.sim_id
        if seed == s + 99999:
            return TapeGenerator("S", self.tape, self.col)
        if seed == s + 50000:
            self.calls_50000 += 1
            if self.calls_50000 == 1:
                return TapeGenerator("A", self.tape, self.col)
            g = TapeGenerator("B", self.tape, self.col, self.reg)
            if self.param_t:
                g = ParamTGenerator("B", self.tape, self.col, self.reg)
            if self.boot is not None:
                g = BootGenerator("B", self.tape, self.col, self.reg, self.boot)
            return g
        # NB: _stable_asset_seed collides for 'SSO' and 'QQQ' (both hash to 486): the reference's
        # two streams then share a seed; on a tape each asset has its own column, served in call order.
        if self.next_asset is not None:
            assert seed == self.eng._stable_asset_seed(s, abi.ASSET_NAMES[self.next_asset])
            return TapeGenerator("T%d" % self.next_asset, self.tape, self.col)
        for a, name in enumerate(abi.ASSET_NAMES):
            if seed == self.eng._stable_asset_seed(s, name) and a not in self.served:
                self.served.add(a)
                return TapeGenerator("T%d" % a, self.tape, self.col)
        raise AssertionError(f"unexpected default_rng({seed})")

    def __enter__(self):
        self.orig = np.random.default_rng
        np.random.default_rng = self.factory
        return self

    def __exit__(self, *a):
        np.random.default_rng = self.orig


class ParamTGenerator(TapeGenerator):
    """stream B of the parametric generator: standard_t(df, size=3) per day (engine.py:91)."""

    def standard_t(self, df, size=None):
        v = self.tp.x[self.g_mvn, :, self.c].copy()
        self.g_mvn += 1
        return v


class BootGenerator(TapeGenerator):
    """stream B of the legacy_hybrid bootstrap (bootstrap.py:310-458, engine.py:50-80): per block a geometric
    length, up to two momentum coins, block indices / sub-section offsets (or, for an empty pool, the Student-t and
    normal draws of a synthetic block, bootstrap.py:295-298); then the (n_days, 3) noise matrix of sample_returns and
    one standard_t(size=3) per day in generate_fat_tailed_returns."""

    def __init__(self, role, tape, col, regime_full, boot):
        super().__init__(role, tape, col, regime_full)
        self.boot = boot
        self.k = -1
        self.cur = 0
        self.in_sampler = True

    def geometric(self, p, size=None):
        assert size is None and abs(p - 1.0 / self.boot.block_mean) < 1e-15
        self.k += 1
        v = int(self.tp.boot_len[self.k, self.c])
        ln = min(int(np.clip(v, self.boot.block_min, self.boot.block_max)), self.tp.gen_days - self.cur)
        regime = int(np.median(self.regime_full[self.cur:self.cur + ln]))
        self.cur += ln
        off = ln < self.boot.block_size
        # the calls this block will make, in order: ('syn', pool) = synthetic block of an empty pool, ('u', slot) = one
        # integer / choice draw served from boot_u[k][slot]
        self.expect = []
        for q, pools, idx_slot, off_slot in ((0, self.boot.pool_a, 1, 2), (1, self.boot.pool_b, 4, 5)):
            self.expect.append(("u", idx_slot) if pools[regime].n_blocks > 0 else ("syn", q))
            if off:
                self.expect.append(("u", off_slot))
        self.pending_normals = 0
        return v

    def _next_u(self):
        kind, slot = self.expect.pop(0)
        assert kind == "u", (kind, slot)
        return float(self.tp.boot_u[self.k, slot, self.c])

    def random(self, size=None, *a, **k):
        assert size is None and self.expect and self.expect[0] in (("u", 1), ("u", 4))   # momentum coin before a block index
        return float(self.tp.boot_u[self.k, 0 if self.expect[0][1] == 1 else 3, self.c])

    def integers(self, low, high=None, *a, **k):
        u = self._next_u()
        return int(low + min(int(np.floor(u * (high - low))), high - low - 1))

    def choice(self, a, size=None, replace=True, p=None, **k):
        assert size is None and p is not None and self.expect[0] == ("u", 4)
        u = self._next_u()
        cdf = np.asarray(p, dtype=float).cumsum()
        cdf /= cdf[-1]
        return int(min(cdf.searchsorted(u, side="right"), a - 1))

    def standard_t(self, df, size=None):
        if isinstance(size, tuple):                  # bootstrap.py:400
            assert size == (self.tp.gen_days, 3)
            self.in_sampler = False
            return self.tp.boot_t[:, :, self.c].copy()
        if self.in_sampler:                          # bootstrap.py:295 synthetic block of an empty pool
            kind, q = self.expect.pop(0)
            assert kind == "syn" and size == self.boot.block_size
            self.pending_normals = 2
            return self.tp.boot_syn[self.k, q, :, self.c].copy()
        v = self.tp.x[self.g_mvn, :, self.c].copy()  # engine.py:68
        self.g_mvn += 1
        return v

    def normal(self, loc=0.0, scale=1.0, size=None):
        if self.in_sampler:                          # bootstrap.py:297-298: VIX / IRX columns of a synthetic block, never read
            assert self.pending_normals > 0 and size == self.boot.block_size
            self.pending_normals -= 1
            return np.zeros(size)
        return super().normal(loc, scale, size)


def import_reference():
    sys.path.insert(0, REF)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
    from letf import config as ref_cfg
    from letf.simulation import engine as ref_engine
    from letf import strategy as ref_strategy
    return ref_cfg, ref_engine, ref_strategy


def apply_ref_cfg(ref_cfg, ns):
    """mirror our config snapshot into the reference's module globals (in-process)."""
    saved = {}
    for k in ("USE_RANDOM_START", "RANDOM_START_METHOD", "RANDOM_START_BUFFER_YEARS", "USE_MOMENT_MATCHING",
              "USE_ANTITHETIC_VARIATES", "SIM_ENGINE_MODE", "STUDENT_T_DF", "USE_BLOCK_BOOTSTRAP", "BOOTSTRAP_WEIGHT",
              "BOOTSTRAP_BLOCK_MIN", "BOOTSTRAP_BLOCK_MAX", "BOOTSTRAP_BLOCK_MEAN", "BOOTSTRAP_BLOCK_SIZE",
              "BOOTSTRAP_MOMENTUM_BIAS_BY_REGIME", "START_REGIME_PROBABILITIES"):
        saved[k] = getattr(ref_cfg, k)
        setattr(ref_cfg, k, getattr(ns, k))
    saved["STRATEGIES"] = dict(ref_cfg.STRATEGIES)
    for k, v in synthetic.EXTRA_BENCHMARKS.items():
        ref_cfg.STRATEGIES[k] = dict(v)
    return saved


def restore_ref_cfg(ref_cfg, saved):
    for k, v in saved.items():
        if k == "STRATEGIES":
            ref_cfg.STRATEGIES.clear(); ref_cfg.STRATEGIES.update(v)
        else:
            setattr(ref_cfg, k, v)


def run_reference_full(case, ref_cfg, ref_engine):
    ns, model, params, keep, tape = case_inputs(case)
    reg_all = tape.regime_full(params)
    saved = apply_ref_cfg(ref_cfg, ns)
    sids = case["strategies"]
    n_trials = case["n_tapes"] * (2 if params.antithetic_pairs else 1)
    exp = {s: {"final_wealth": [], "max_dd": [], "num_trades": [], "journal_count": [], "journal_volume": []}
           for s in sids}
    meta = {"start_offset": [], "regime_days": [], "jump_days": [], "repaired": []}
    ref_model = {k: v for k, v in model.items() if k != "synthetic_normal"}
    boot = None
    if params.market_mode == abi.MARKET_BOOTSTRAP:
        # the reference's own sampler over the same frame (bootstrap.py:27-47; its pickle cache is bypassed)
        import contextlib, io
        from letf.simulation.bootstrap import BlockBootstrapReturns
        with contextlib.redirect_stdout(io.StringIO()):
            ref_model["bootstrap_sampler"] = BlockBootstrapReturns(model["bootstrap_sampler"].df,
                                                                   block_size=ns.BOOTSTRAP_BLOCK_SIZE)
        boot = params.boot
    try:
        for k in range(n_trials):
            col = k >> 1 if params.antithetic_pairs else k
            anti = bool(k & 1) if params.antithetic_pairs else False
            sim_id = 1000 + col
            with PatchedRng(tape, col, sim_id, reg_all[:, col], ref_engine,
                            param_t=params.market_mode == abi.MARKET_PARAM_T, boot=boot):
                res = ref_engine.simulate_single_path_fixed((sim_id, case["years"], ref_model, {}, sids, anti))
            for s in sids:
                r = res[s]
                assert "Error" not in r, r.get("Error")
                exp[s]["final_wealth"].append(float(r["Final_Wealth"]))
                exp[s]["max_dd"].append(float(r["Max_DD"]))
                exp[s]["num_trades"].append(int(r["Num_Trades"]))
                tl = r.get("Trade_List", [])
                exp[s]["journal_count"].append(len(tl))
                exp[s]["journal_volume"].append(float(sum(t["dollar_amount"] for t in tl)))
            md = res[sids[0]]["Metadata"]
            rp = np.asarray(res[sids[0]]["Regime_Path"])
            assert np.array_equal(rp, reg_all[tape.start_offset[col]:tape.start_offset[col] + params.sim_days, col]) \
                or not params.random_start
            meta["start_offset"].append(int(tape.start_offset[col]) if params.random_start else 0)
            meta["regime_days"].append([int((rp == 0).sum()), int((rp == 1).sum())])
            meta["jump_days"].append(int(md.get("stress_summary", {}).get("jump_days", 0)))
            meta["repaired"].append(bool(md["layer_checks"].get("repaired_after_validation", False)))
    finally:
        restore_ref_cfg(ref_cfg, saved)
    return exp, meta


def run_reference_layers(case, ref_cfg, ref_engine, ref_strategy):
    """NORMAL-market shapes (configs 1/2/5): the reference has no switch for these, so the layer
    functions are called directly, as diagnose.py:155-166 does: simulate_etf_returns_from_layers
    (engine.py:360) per asset on the tape's correlated normal underlyings, then the benchmark
    rule run_strategy_fixed (strategy.py:85-95) and the statistics of engine.py:578-586."""
    import pandas as pd
    ns, model, params, keep, tape = case_inputs(case)
    saved = apply_ref_cfg(ref_cfg, ns)
    sids = case["strategies"]
    T = params.sim_days
    exp = {s: {"final_wealth": [], "max_dd": [], "num_trades": [], "journal_count": [], "journal_volume": []}
           for s in sids}
    meta = {"start_offset": [], "regime_days": [], "jump_days": [], "repaired": []}
    sn = model["synthetic_normal"]
    ch = np.asarray(sn["chol"]) if "chol" in sn else np.linalg.cholesky(np.asarray(sn["cov"]))
    mu = np.asarray(sn["mu"], dtype=float)
    try:
        for col in range(case["n_tapes"]):
            x = tape.x[:, :, col]
            spy = mu[0] + ch[0, 0] * x[:, 0]
            qqq = mu[1] + (ch[1, 0] * x[:, 0] + ch[1, 1] * x[:, 1])
            tlt = mu[2] + (ch[2, 0] * x[:, 0] + ch[2, 1] * x[:, 1] + ch[2, 2] * x[:, 2])
            regime = np.zeros(T, dtype=int)
            vix = np.full(T, sn["vix"]); rf = np.full(T, sn["rf"])
            sim_id = 2000 + col
            cols = {}
            for a, name in enumerate(abi.ASSET_NAMES):
                if not (params.asset_mask >> a) & 1:
                    continue
                conf = ns.LETF_CONFIGS[name]
                und = ref_engine.map_underlying_series_for_asset(name, conf.get("beta_to_spy", 1.0), spy, qqq, tlt)
                with PatchedRng(tape, col, sim_id, None, ref_engine) as pr:
                    pr.next_asset = a
                    cols[f"{name}_Ret"] = ref_engine.simulate_etf_returns_from_layers(
                        name, conf, und, regime, vix, rf, sim_id, None, None, None)
            df = pd.DataFrame(cols)
            df["Cash_Ret"] = ref_cfg.CASH_RATE_BY_REGIME[0] / 252
            for name in list(cols):
                a = name[:-4]
                df[f"{a}_Price"] = (1 + df[name].fillna(0)).cumprod() * 100
            df["VIX"] = vix
            for s in sids:
                ec, nt = ref_strategy.run_strategy_fixed(df, s, regime, {}, apply_costs=True, trade_journal=None)
                e = ec.values
                rm = np.maximum.accumulate(e)
                dd = (e - rm) / np.where(rm > 0, rm, 1.0)
                exp[s]["final_wealth"].append(float(e[-1])); exp[s]["max_dd"].append(float(abs(dd.min())))
                exp[s]["num_trades"].append(int(nt)); exp[s]["journal_count"].append(0)
                exp[s]["journal_volume"].append(0.0)
            meta["start_offset"].append(0); meta["regime_days"].append([T, 0]); meta["jump_days"].append(0)
            meta["repaired"].append(False)
        if case["name"] == "voldrag_c1":
            # validation.py:45-57 arithmetic on the same shocks: prod(1 + 3 * r) - 1
            kat = []
            for col in range(case["n_tapes"]):
                r = 0 + (0.15 / np.sqrt(252)) * tape.x[:, 0, col]   # rng.normal(0, daily_std, n_days)
                kat.append([float(np.prod(1 + L * r) - 1) for L in (3.0, 2.0, 1.0)])
            meta["validation_py_annual_return"] = kat
    finally:
        restore_ref_cfg(ref_cfg, saved)
    return exp, meta


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--outdir", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    ref_cfg, ref_engine, ref_strategy = import_reference()
    import pandas, scipy, numba
    versions = {"numpy": np.__version__, "pandas": pandas.__version__, "scipy": scipy.__version__,
                "numba": numba.__version__, "python": sys.version.split()[0]}
    os.makedirs(args.outdir, exist_ok=True)
    for case in CASES:
        if args.only and case["name"] != args.only:
            continue
        t0 = time.time()
        if case["variant"].startswith("normal_"):
            exp, meta = run_reference_layers(case, ref_cfg, ref_engine, ref_strategy)
        else:
            exp, meta = run_reference_full(case, ref_cfg, ref_engine)
        doc = {"case": case, "versions": versions, "expected": exp, "meta": meta,
               "generator": "oracle/make_golden.py (live reference at /root/reference, TapeGenerator)"}
        with open(os.path.join(args.outdir, case["name"] + ".json"), "w") as f:
            json.dump(doc, f, indent=0)
        print(f"{case['name']}: {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
