"""Parity of the CUDA path (through the C-ABI of include/letf_b200.h) with the oracle and with
the golden vectors generated from the live reference.  Runs on the B200 box (-m gpu).

Bars (north star): shared pre-drawn shocks -> regime-state sequences (regime day counts, start
offsets, jump days) and trade / journal counts bit-exact, terminal wealth <= 1e-6 relative with
fp64 accumulation (we hold 1e-11); fp32 throughput mode has its own stated tolerance; on-device
Philox draws must match the reference's distributions (KS) and summary quantiles.
"""
import numpy as np
import pytest
from scipy import stats

from letf_monte_carlo_b200 import abi, synthetic
from letf_monte_carlo_b200.params import build_params
from letf_monte_carlo_b200.tape import HostTape, make_tape
from oracle import oracle as orc
from tests.common import GOLDEN_NAMES, assert_matches_reference, case_inputs, expected_arrays, load_case

pytestmark = pytest.mark.gpu
ALL19 = ["S%d" % i for i in range(1, 20)]


def _engine(params, keep, precision):
    from letf_monte_carlo_b200.engine import PathEngine
    return PathEngine(params, keep, precision=precision, device=0)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_f64_tape_matches_reference_golden(name, engine_lib):
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    n = tape.n_tapes * (2 if params.antithetic_pairs else 1)
    with _engine(params, keep, "f64") as eng:
        out = eng.run_host(n, rng_mode=abi.RNG_TAPE, tape=tape, journal=True)
    assert_matches_reference(out, doc, rtol=1e-11, dd_atol=1e-12)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_f64_tape_matches_oracle_all_outputs(name, engine_lib):
    """device buffers (letf_run) instead of the host convenience path, compared field by field"""
    import torch
    from letf_monte_carlo_b200.engine import DeviceTape
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    n = tape.n_tapes * (2 if params.antithetic_pairs else 1)
    o = orc.run(params, tape)
    with _engine(params, keep, "f64") as eng:
        dt = DeviceTape(tape, eng.device)
        out = eng.run(n, rng_mode=abi.RNG_TAPE, tape=dt, journal=True)
        torch.cuda.synchronize()
        h = out.to_host()
    np.testing.assert_allclose(h.final_wealth, o.final_wealth, rtol=1e-11)
    np.testing.assert_allclose(h.max_dd, o.max_dd, atol=1e-12)
    for k in ("num_trades", "journal_count", "start_offset", "flags", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(h, k), getattr(o, k), err_msg=k)
    np.testing.assert_allclose(h.journal_volume, o.journal_volume, rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize("name", ["inst_1y", "layers_c2_5y", "param_t_1y", "vix_te_1y", "voldrag_c1", "boot_1y"])
def test_f32_tape_close_to_reference(name, engine_lib):
    """fp32 throughput mode on the same tape: 2e-4 relative on terminal wealth over <= 5 years,
    1e-4 absolute on drawdown, < 2 % of (strategy, trial) cells with a different trade count."""
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    n = tape.n_tapes * (2 if params.antithetic_pairs else 1)
    with _engine(params, keep, "f32") as eng:
        out = eng.run_host(n, rng_mode=abi.RNG_TAPE, tape=tape, journal=True)
    assert_matches_reference(out, doc, rtol=2e-4, dd_atol=1e-4, exact_counts=False)
    assert np.mean(out.num_trades != expected_arrays(doc)[2]) < 0.02


@pytest.mark.parametrize("name", ["inst_10y", "inst_30y", "boot_10y", "layers_c2_30y"])
def test_f32_tape_close_to_reference_at_benchmark_horizons(name, engine_lib):
    """fp32 throughput mode at the horizons bench.py runs (10 and 30 years = 7 560 days), same pre-drawn tape as the
    live-reference golden.  Stated fp32 tolerance at 30 years: terminal wealth 5e-4 relative (measured <= 7e-5), max
    drawdown 5e-5 absolute, trade counts off by at most 2 in at most 3 of the <= 92 (strategy, trial) cells of a case,
    integer path state identical."""
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    n = tape.n_tapes * (2 if params.antithetic_pairs else 1)
    with _engine(params, keep, "f32") as eng:
        out = eng.run_host(n, rng_mode=abi.RNG_TAPE, tape=tape, journal=True)
    assert_matches_reference(out, doc, rtol=5e-4, dd_atol=5e-5, exact_counts=False)
    fw, dd, nt, jc, jv = expected_arrays(doc)
    assert np.sum(out.num_trades != nt) <= 3 and np.abs(out.num_trades - nt).max() <= 2
    print("f32 tape", name, "max rel", float(np.max(np.abs(out.final_wealth / fw - 1))), "dd", float(np.max(np.abs(out.max_dd - dd))),
          "trade-cell mismatch", float(np.mean(out.num_trades != nt)))


def test_constant_window_corner(engine_lib):
    from tests.test_kernel_emulation import drifty_case
    params, keep, tape = drifty_case(years=3, n=64, seed=11)
    o = orc.run(params, tape)
    with _engine(params, keep, "f64") as eng:
        out = eng.run_host(64, rng_mode=abi.RNG_TAPE, tape=tape, journal=True)
    np.testing.assert_allclose(out.final_wealth, o.final_wealth, rtol=1e-11)
    np.testing.assert_array_equal(out.num_trades, o.num_trades)
    np.testing.assert_array_equal(out.journal_count, o.journal_count)


@pytest.mark.parametrize("variant,years,precision", [("default", 2, "f64"), ("vix_te", 1, "f64"), ("no_joint", 1, "f64"),
                                                     ("durations", 1, "f64"), ("normal_c2", 3, "f64"),
                                                     ("boot", 2, "f64"), ("boot_early", 1, "f64")])
def test_philox_record_replays_through_oracle(variant, years, precision, engine_lib):
    """the on-device Philox path records the variates it consumed; the oracle on that tape must agree"""
    sids = ALL19 if variant != "normal_c2" else ["S1", "S3", "B_UPRO", "S4", "S6", "S19"]
    doc = {"case": dict(variant=variant, years=years, n_tapes=40, seed=0,
                        cfg={"SIM_ENGINE_MODE": "legacy_hybrid"} if variant.startswith("boot") else {}, strategies=sids)}
    ns, model, params, keep, _ = case_inputs(doc)
    n_t = 40
    n = n_t * (2 if params.antithetic_pairs else 1)
    G = params.sim_days + params.buffer_days
    from letf_monte_carlo_b200.tape import boot_tape_shape
    rec = HostTape(n_t, G, params.sim_days, 512, *boot_tape_shape(params))
    with _engine(params, keep, precision) as eng:
        out = eng.run_host(n, seed=4242, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec, journal=True)
        plain = eng.run_host(n, seed=4242, rng_mode=abi.RNG_PHILOX, journal=True)
    assert not (out.flags & abi.FLAG_SPELL_OVERFLOW).any()
    o = orc.run(params, rec)
    np.testing.assert_allclose(out.final_wealth, o.final_wealth, rtol=1e-10)
    np.testing.assert_array_equal(out.num_trades, o.num_trades)
    np.testing.assert_array_equal(out.journal_count, o.journal_count)
    np.testing.assert_array_equal(out.regime_days, o.regime_days)
    np.testing.assert_array_equal(out.jump_days, o.jump_days)
    np.testing.assert_array_equal(out.flags, o.flags)
    np.testing.assert_array_equal(plain.final_wealth, out.final_wealth)
    np.testing.assert_array_equal(plain.num_trades, out.num_trades)


def test_partition_invariance_and_kernel_variants(engine_lib):
    """results are a pure function of (seed, trial id): independent of the launch partition and of
    which instantiation (lean benchmark build vs the all-strategies build) runs them"""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    model = synthetic.fallback_regime_model("default")
    p_lean, k1 = build_params(["S1", "S2", "S3"], 1, model, ns)
    p_full, k2 = build_params(["S1", "S2", "S3", "S4", "S19"], 1, model, ns)
    p_full.asset_mask = p_lean.asset_mask
    n = 3000
    with _engine(p_lean, k1, "f32") as e1, _engine(p_full, k2, "f32") as e2:
        a = e1.run_host(n, seed=7)
        b0 = e1.run_host(1000, seed=7, first_trial=0)
        b1 = e1.run_host(2000, seed=7, first_trial=1000)
        c = e2.run_host(n, seed=7)
        assert e1.launch_info()["kernel"] != e2.launch_info()["kernel"]      # (pair kernel vs routed split pipeline)
    np.testing.assert_array_equal(a.final_wealth[:, :1000], b0.final_wealth)
    np.testing.assert_array_equal(a.final_wealth[:, 1000:], b1.final_wealth)
    np.testing.assert_array_equal(a.max_dd[:, 1000:], b1.max_dd)
    # same shocks through a different instantiation: tiny fp32 reassociation differences only
    np.testing.assert_allclose(c.final_wealth[:3], a.final_wealth, rtol=1e-4)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_philox_draw_distributions_on_device(precision, engine_lib):
    doc = {"case": dict(variant="vix_te", years=2, n_tapes=1, seed=0, cfg={}, strategies=["S1", "S2"])}
    ns, model, params, keep, _ = case_inputs(doc)
    params.asset_mask = 0x3F
    n_t = 64
    G, T = params.sim_days + params.buffer_days, params.sim_days
    rec = HostTape(n_t, G, T, 512)
    with _engine(params, keep, precision) as eng:
        eng.run_host(2 * n_t, seed=99, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
    reg = rec.regime_full(params)

    def ks(sample, dist, alpha=1e-4):
        d, p = stats.kstest(sample, dist.cdf)
        assert p > alpha, (d, p)

    ks(rec.x.ravel(), stats.norm)
    used = np.arange(G)[:, None] < (rec.start_offset[None, :] + T)
    used[0] = False
    ks(rec.vix_n[used], stats.norm)
    ks(rec.irx_n[used], stats.norm)
    for r in (0, 1):
        ks(rec.chi2[reg == r], stats.chi2(params.joint[r].nu))
    ks(rec.liq_n[1:].ravel(), stats.norm)
    ks(rec.jump_u[1:].ravel(), stats.uniform)
    ks(rec.jump_t[1:].ravel(), stats.t(5))
    for a in range(abi.N_ASSETS):
        ks(rec.te_t[a, 1:].ravel(), stats.t(params.assets[a].te_df))


def test_config1_vol_drag_known_answer(engine_lib):
    """validation.py:18-87 (BASELINE config 1): 10 000 paths x 252 days, zero drift, sigma 15 %, L = 3:
    median annual return within 1.5 % (abs) of -0.5 L^2 sigma^2 = -10.125 %; and L = 2 (validation.py:90-188)
    within 2 % of -4.5 %."""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                    USE_ANTITHETIC_VARIATES=False, LETF_CONFIGS=synthetic.pure_leverage_assets())
    ns.ASSETS = ns.LETF_CONFIGS
    params, keep = build_params(["B_UPRO", "S3", "S2"], 1, synthetic.normal_regime_model("c1"), ns)
    for precision in ("f32", "f64"):
        with _engine(params, keep, precision) as eng:
            out = eng.run_host(10000, seed=42)
        ann = out.final_wealth / params.initial_capital - 1.0
        assert abs(np.median(ann[0]) - (-0.10125)) < 0.015
        assert abs(np.median(ann[1]) - (-0.045)) < 0.02
        assert abs(np.median(ann[2]) - (-0.01125)) < 0.02
        assert (out.flags == 0).all()


def test_quantiles_match_reference_distribution(engine_lib):
    """summary-quantile agreement of Philox runs with the reference's arithmetic on numpy draws
    (validation.py:45-57 vectorised): two-sample KS on terminal wealth, 1e5 paths each"""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                    USE_ANTITHETIC_VARIATES=False, LETF_CONFIGS=synthetic.pure_leverage_assets())
    ns.ASSETS = ns.LETF_CONFIGS
    params, keep = build_params(["B_UPRO"], 1, synthetic.normal_regime_model("c1"), ns)
    with _engine(params, keep, "f32") as eng:
        out = eng.run_host(100000, seed=1)
    rng = np.random.default_rng(42)
    r = rng.normal(0, 0.15 / np.sqrt(252), (100000, 252))
    ref = 1e4 * np.prod(1 + 3.0 * r, axis=1)
    d, p = stats.ks_2samp(out.final_wealth[0], ref)
    assert p > 1e-3, (d, p)
    for q in (10, 25, 50, 75, 90):
        a, b = np.percentile(out.final_wealth[0], q), np.percentile(ref, q)
        assert abs(a / b - 1) < 0.01


def test_run_strategies_matches_oracle(engine_lib):
    """run_strategy_fixed seam: strategy rules on caller-supplied series (strategy.py:55)"""
    import ctypes as C
    rng = np.random.default_rng(3)
    T, n = 600, 24
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(ALL19, T / 252, synthetic.fallback_regime_model(), ns)
    rets = {a: np.ascontiguousarray(rng.standard_t(4, (T, n)) * s) for a, s in
            zip(abi.ASSET_NAMES, (0.03, 0.03, 0.02, 0.025, 0.01, 0.013))}
    vix = np.ascontiguousarray(15 + 10 * rng.random((T, n)))
    regime = np.ascontiguousarray((rng.random((T, n)) < 0.2).astype(np.int32))
    tlt = np.ascontiguousarray(rng.normal(0, 0.008, (T, n)))
    from oracle.oracle import HostOutputs
    o = HostOutputs(params.n_strategies, n)
    s = abi.Series()
    s.n_paths, s.sim_days = n, T
    for a, name in enumerate(abi.ASSET_NAMES):
        s.asset_ret[a] = rets[name].ctypes.data
    s.tlt_ret, s.vix, s.regime = tlt.ctypes.data, vix.ctypes.data, regime.ctypes.data
    assert orc.lib().letf_oracle_run_strategies(C.byref(params), C.byref(s), C.byref(o.as_struct())) == 0
    with _engine(params, keep, "f64") as eng:
        h = eng.run_strategies(rets, tlt_ret=tlt, vix=vix, regime=regime, journal=True).to_host()
    np.testing.assert_allclose(h.final_wealth, o.final_wealth, rtol=1e-11)
    np.testing.assert_allclose(h.max_dd, o.max_dd, atol=1e-12)
    np.testing.assert_array_equal(h.num_trades, o.num_trades)
    np.testing.assert_array_equal(h.journal_count, o.journal_count)


def test_trade_list_replay_matches_oracle(engine_lib):
    doc = load_case("inst_1y")
    ns, model, params, keep, tape = case_inputs(doc)
    sids = doc["case"]["strategies"]
    with _engine(params, keep, "f64") as eng:
        for sid in ("S7", "S8", "S13", "S19"):
            s = sids.index(sid)
            got = eng.trade_list(3, s, rng_mode=abi.RNG_TAPE, tape=tape)
            ref = orc.trade_list(params, tape, 3, s)
            assert len(got) == len(ref) == doc["expected"][sid]["journal_count"][3]
            for g, r in zip(got, ref):
                assert (g["day_index"], g["asset"], g["action"]) == (r["day_index"], r["asset"], r["action"])
                assert abs(g["shares"] / r["shares"] - 1) < 1e-10 and abs(g["price"] / r["price"] - 1) < 1e-10
        assert eng.trade_list(3, sids.index("S4"), rng_mode=abi.RNG_TAPE, tape=tape) == []   # not in TAXABLE_IDS


def test_reduce_matches_numpy(engine_lib):
    import torch
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(["S1", "S2", "S3", "S4"], 1, synthetic.fallback_regime_model(), ns)
    n = 20000
    with _engine(params, keep, "f32") as eng:
        out = eng.run(n, seed=3)
        sm = eng.new_summary(bench_strategy=1)
        eng.reduce(out, sm, n_trials=n // 2)
        eng.reduce(out_slice(out, n // 2, n), sm)          # accumulates
        torch.cuda.synchronize()
        h = out.to_host()
        hist = sm.hist.cpu().numpy(); mom = sm.moments.cpu().numpy(); ext = sm.extrema.cpu().numpy()
        cnt = sm.counters.cpu().numpy()
        eng.reset_summary(sm)
        torch.cuda.synchronize()
        assert int(sm.hist.sum()) == 0 and int(sm.counters.sum()) == 0
    W0, years = params.initial_capital, params.sim_years
    for s in range(4):
        w = h.final_wealth[s]
        # a switch on a -99.9 % day can push S4's equity below zero (ret = r - cost, strategy.py:140-150);
        # the reduction clamps such ratios into the first bin (letf_reduce.cu)
        ratio = w / W0
        metrics = [np.log10(np.maximum(ratio, 1e-300)), np.maximum(ratio, 0.0) ** (1 / years) - 1, h.max_dd[s]]
        for q, m in enumerate(metrics):
            ref, _ = np.histogram(np.clip(m, sm.lo[q], np.nextafter(sm.hi[q], -np.inf)), bins=sm.bin_edges(q))
            assert np.abs(hist[s, q] - ref).sum() <= 4          # bin-edge rounding only
            assert hist[s, q].sum() == n
            np.testing.assert_allclose(mom[s, q, 0], m.sum(), rtol=1e-9)
            np.testing.assert_allclose(mom[s, q, 1], (m * m).sum(), rtol=1e-9)
            np.testing.assert_allclose(ext[s, q], [m.min(), m.max()], rtol=1e-12)
        assert cnt[s, 0] == n
        assert cnt[s, 1] == int((w > h.final_wealth[1]).sum())       # wins vs S2 by trial index
        assert cnt[s, 2] == int(h.num_trades[s].sum())


def out_slice(out, lo, hi):
    """view of columns [lo, hi) of a DeviceOutputs as a new struct-compatible object"""
    from letf_monte_carlo_b200.outputs import DeviceOutputs, FIELDS
    v = DeviceOutputs.__new__(DeviceOutputs)
    v.n_strategies = out.n_strategies
    v.n = hi - lo
    for k in FIELDS:
        t = getattr(out, k)
        setattr(v, k, None if t is None else t[..., lo:hi].contiguous())
    return v


def test_errors_are_reported_not_swallowed(engine_lib):
    from letf_monte_carlo_b200.engine import EngineError, NonFiniteTrials, PathEngine
    ns = synthetic.config_namespace()
    params, keep = build_params(["S1"], 1, synthetic.fallback_regime_model(), ns)
    with _engine(params, keep, "f32") as eng:
        with pytest.raises(EngineError):
            eng.run_host(4, rng_mode=abi.RNG_TAPE, tape=None)            # tape mode without a tape
        small = HostTape(1, 10, 5, 4)
        with pytest.raises(EngineError):
            eng.run_host(2, rng_mode=abi.RNG_TAPE, tape=small)           # wrong day counts
    bad = build_params(["S4"], 1, synthetic.fallback_regime_model(), ns)[0]
    bad.strategies[0].p[0] = 150.0                                       # unsupported SMA window
    with pytest.raises(EngineError):
        PathEngine(bad, (), precision="f32", device=0)
    # non-finite generated returns -> the reference raises ValueError out of the trial
    nf, k = build_params(["S2"], 1, synthetic.fallback_regime_model(), ns)
    nf.joint[0].mu[0] = float("nan")
    with _engine(nf, k, "f32") as eng:
        with pytest.raises(NonFiniteTrials):
            eng.run_host(8, seed=1)


def test_full_size_properties_config2(engine_lib):
    """BASELINE config 2 at its full size (100k trials x 7 560 days, TQQQ/UPRO/SSO with expense,
    borrow and tracking error): size-independent properties."""
    import torch
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                    USE_ANTITHETIC_VARIATES=False)
    params, keep = build_params(["S1", "B_UPRO", "S3"], 30, synthetic.normal_regime_model("c2"), ns)
    n = 100000
    with _engine(params, keep, "f32") as eng:
        out = eng.run(n, seed=2024)
        halves = [eng.run(n // 2, seed=2024, first_trial=f) for f in (0, n // 2)]
        torch.cuda.synchronize()
        h = out.to_host()
        h0, h1 = (x.to_host() for x in halves)
    np.testing.assert_array_equal(h.final_wealth, np.concatenate([h0.final_wealth, h1.final_wealth], axis=1))
    assert (h.flags == 0).all() and np.isfinite(h.final_wealth).all() and (h.final_wealth > 0).all()
    assert (h.max_dd >= 0).all() and (h.max_dd <= 1).all() and (h.num_trades == 0).all()
    # E[log growth] per day = L mu - L^2 sigma^2 / 2 - ER/252 - borrow + O(mu^2): known answer within 5 se
    T = params.sim_days
    vol = np.array([0.16, 0.24]) / np.sqrt(252); mu = np.array([0.08, 0.10]) / 252
    for s, (a, u) in enumerate((("TQQQ", 1), ("UPRO", 0), ("SSO", 0))):
        c = ns.ASSETS[a]; L = c["leverage"]
        cost = c["expense_ratio"] / 252 + (L - 1) * (0.045 + c["borrow_spread"]) / 252
        m = L * mu[u] - cost
        g = m - 0.5 * (L * vol[u]) ** 2 - 0.5 * m * m            # second-order log expansion
        lg = np.log(h.final_wealth[s] / params.initial_capital) / T
        se = L * vol[u] / np.sqrt(T * n)
        assert abs(lg.mean() - g) < 5 * se + 2e-6, (a, lg.mean(), g, se)


@pytest.mark.parametrize("years", [5, 30])
def test_lean_fast_kernel_vs_generic_and_oracle(years, engine_lib):
    """bench.py's c2 workload runs the lean buy&hold build with the loop-unswitched FAST day step.
    Chain of evidence on the device: FAST (Philox) == generic fp32 build recording the same counters
    (fp32 reassociation only), and the oracle replaying that recorded tape agrees with both - at 5 years and at the
    benchmarked 30-year horizon (stated fp32 tolerance there: 1e-3 relative, measured value printed)."""
    sids = ["S1", "B_UPRO", "S3"]
    n = 256 if years == 5 else 64
    doc = {"case": dict(variant="normal_c2", years=years, n_tapes=n, seed=0, cfg={}, strategies=sids)}
    ns, model, params, keep, _ = case_inputs(doc)
    rec = HostTape(n, params.sim_days, params.sim_days, 1)
    with _engine(params, keep, "f32") as eng:
        lean = eng.run_host(n, seed=31337)
        regs_lean = eng.launch_info()["regs_per_thread"]
        gen = eng.run_host(n, seed=31337, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
        regs_gen = eng.launch_info()["regs_per_thread"]
    assert regs_lean < regs_gen                      # two different instantiations really ran
    o = orc.run(params, rec)
    k = 1 if years == 5 else 5
    np.testing.assert_allclose(lean.final_wealth, gen.final_wealth, rtol=5e-5 * k)
    np.testing.assert_allclose(lean.max_dd, gen.max_dd, atol=5e-6 * k)
    np.testing.assert_allclose(gen.final_wealth, o.final_wealth, rtol=2e-4 * k)
    np.testing.assert_allclose(lean.final_wealth, o.final_wealth, rtol=2e-4 * k)
    np.testing.assert_allclose(lean.max_dd, o.max_dd, atol=1e-4)
    assert (lean.flags == 0).all() and (o.flags == 0).all()
    print("c2 FAST", years, "y: max rel vs oracle", float(np.max(np.abs(lean.final_wealth / o.final_wealth - 1))),
          "vs generic", float(np.max(np.abs(lean.final_wealth / gen.final_wealth - 1))))


@pytest.mark.parametrize("sids,years", [(["S2", "B_QQQ", "B_TLT"], 2), (["S1", "S2", "S3"], 1), (["S2", "B_QQQ", "B_TLT"], 30)],
                         ids=["c3_2y", "letf3_1y", "c3_30y"])
def test_joint_pair_kernel_vs_generic_and_oracle(sids, years, engine_lib, monkeypatch):
    """bench.py's c3 workload (BASELINE configs[2]) runs the antithetic-pair kernel (letf_jfast.cuh): one thread carries both
    twins of a pair and draws their shared Philox blocks once.  Chain of evidence on the device: pair kernel (Philox) ==
    the lean build's general loop on the same counters == the generic fp32 build recording those counters, and the
    fp64 oracle replaying the recorded tape agrees with all of them - at 30 years too (the benchmarked horizon)."""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, years, synthetic.fallback_regime_model("default"), ns)
    n_t = 96 if years < 30 else 24
    n = 2 * n_t
    rec = HostTape(n_t, params.sim_days + params.buffer_days, params.sim_days, 1024)
    with _engine(params, keep, "f32") as eng:
        pair = eng.run_host(n, seed=4242)
        assert eng.launch_info()["kernel"].startswith("f32_joint_t_pair")
        part = eng.run_host(9, seed=4242, first_trial=3)           # starts on an odd twin, ends on an even one
        gen = eng.run_host(n, seed=4242, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
        assert not eng.launch_info()["kernel"].startswith("f32_joint_t_pair")
    monkeypatch.setenv("LETF_NO_JPAIR", "1")
    with _engine(params, keep, "f32") as eng:
        lean = eng.run_host(n, seed=4242)
        assert eng.launch_info()["kernel"] == "f32_joint_t_bench"
    o = orc.run(params, rec)
    tol = 5e-5 if years < 30 else 5e-4          # fp32 reassociation between builds, compounded over the horizon
    for other in (lean, gen):
        np.testing.assert_allclose(pair.final_wealth, other.final_wealth, rtol=tol)
        np.testing.assert_allclose(pair.max_dd, other.max_dd, atol=2e-5)
        np.testing.assert_array_equal(pair.regime_days, other.regime_days)
        np.testing.assert_array_equal(pair.start_offset, other.start_offset)
        np.testing.assert_array_equal(pair.jump_days, other.jump_days)
        np.testing.assert_array_equal(pair.flags, other.flags)
    np.testing.assert_array_equal(part.final_wealth, pair.final_wealth[:, 3:12])
    # fp32 device arithmetic vs the fp64 oracle on identical shocks: stated fp32 tolerance (2e-4 over <= 5 years, 2e-3 at 30 years)
    np.testing.assert_allclose(pair.final_wealth, o.final_wealth, rtol=2e-4 if years < 30 else 2e-3)
    np.testing.assert_allclose(pair.max_dd, o.max_dd, atol=1e-4 if years < 30 else 5e-4)
    np.testing.assert_array_equal(pair.regime_days, o.regime_days)
    np.testing.assert_array_equal(pair.jump_days, o.jump_days)
    np.testing.assert_array_equal(pair.start_offset, o.start_offset)
    np.testing.assert_array_equal((pair.flags & abi.FLAG_REPAIRED) != 0, (o.flags & abi.FLAG_REPAIRED) != 0)
    print("joint_pair", sids, years, "max rel vs oracle", float(np.max(np.abs(pair.final_wealth / o.final_wealth - 1))),
          "vs lean", float(np.max(np.abs(pair.final_wealth / lean.final_wealth - 1))))


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_philox_distribution_matches_live_reference(precision, engine_lib):
    """KS + summary-quantile agreement of on-device Philox runs with the LIVE reference's own NumPy streams
    (tests/golden/ref_distribution_2y.npz: 2 400 trials, institutional_v1 defaults, 2 years, 10 strategies)."""
    from tests.common import assert_same_distribution, reference_distribution
    ref = reference_distribution()
    sids = [str(s) for s in ref["strategies"]]
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, float(ref["years"]), synthetic.fallback_regime_model("default"), ns)
    with _engine(params, keep, precision) as eng:
        out = eng.run_host(40000, seed=2718)
    assert_same_distribution(ref, out, sids)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_philox_distribution_matches_live_reference_at_10_years(precision, engine_lib):
    """tests/golden/ref_distribution_10y.npz: 2 000 trials x 10 years x {S1 S2 S3 S4 S5 S6 S7 S19 QQQ TLT} of the unmodified
    reference with its own NumPy streams (oracle/make_reference_distribution.py --pairs 1000 --years 10).  On-device Philox
    runs, fp32 (the split strategy pipeline) and fp64: KS p > 0.01 per strategy on wealth and drawdown, the reference's
    P10 ... P90 within +- 3 rank points, mean trade counts within 3 %."""
    from tests.common import assert_same_distribution, reference_distribution
    ref = reference_distribution("ref_distribution_10y.npz")
    sids = [str(s) for s in ref["strategies"]]
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, float(ref["years"]), synthetic.fallback_regime_model("default"), ns)
    with _engine(params, keep, precision) as eng:
        out = eng.run_host(40000, seed=314)
    assert_same_distribution(ref, out, sids, p_min=0.01, rank_tol=3.0)


@pytest.mark.parametrize("variant", ["default", "vix_te"], ids=["default_te", "calibrated_te"])
@pytest.mark.parametrize("sids", [["S1", "S2", "S4", "S5", "S6", "S19"], ALL19], ids=["c4_list", "sweep19_list"])
def test_typed_slot_kernels_vs_generic_and_oracle(sids, variant, engine_lib):
    """The typed-slot builds (one compile-time rule per slot; BASELINE config 4 and the reference's default
    19-strategy sweep) against the generic fp32 build recording the same Philox counters, and the oracle
    replaying that tape.  `vix_te`: an explicit tracking_residual_model (non-integer dof, tanh cap) - served by the
    typed builds that read the tracking-error dof at run time, not by the generic 32-slot build."""
    ns = synthetic.config_namespace()
    params, keep = build_params(sids, 2, synthetic.fallback_regime_model(variant), ns)
    n_t = 128
    rec = HostTape(n_t, params.sim_days + params.buffer_days, params.sim_days, 512)
    with _engine(params, keep, "f32") as eng:
        typed = eng.run_host(2 * n_t, seed=99)
        regs_typed = eng.launch_info()["regs_per_thread"]
        gen = eng.run_host(2 * n_t, seed=99, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
    o = orc.run(params, rec)
    # same source expressions, different instantiation: fp32 contraction differences only
    np.testing.assert_allclose(typed.final_wealth, gen.final_wealth, rtol=2e-3)
    assert np.mean(typed.num_trades != gen.num_trades) < 0.01
    np.testing.assert_array_equal(typed.regime_days, gen.regime_days)
    np.testing.assert_array_equal(typed.start_offset, o.start_offset)
    # fp32 signals vs the fp64 oracle on identical shocks: wealth within 1 % except the skew-driven rules
    rel = np.abs(typed.final_wealth - o.final_wealth) / np.abs(o.final_wealth)
    assert np.median(rel) < 1e-4 and np.mean(rel < 1e-2) > 0.97, (np.median(rel), np.mean(rel < 1e-2))
    assert np.mean(typed.num_trades != o.num_trades) < 0.03
    assert regs_typed > 0 and regs_typed != 227        # 227 = the generic f32_joint_t_all build: a typed build ran


@pytest.mark.parametrize("sids", [["S1", "S2", "S4", "S5", "S6", "S19"], ALL19], ids=["c4_list", "sweep19_list"])
def test_split_pipeline_vs_fused_generic_and_oracle(sids, engine_lib, monkeypatch):
    """The split strategy pipeline (pair producer -> per-day records in HBM -> typed rule kernel) is what runs for the c4 and
    default-sweep shapes.  On the device: split == fused typed build (LETF_NO_SPLIT=1) on the same Philox counters to fp32
    reassociation, identical path metadata; and the fp64 oracle replaying the tape the generic build recorded agrees."""
    ns = synthetic.config_namespace()
    params, keep = build_params(sids, 2, synthetic.fallback_regime_model("default"), ns)
    n_t = 160
    rec = HostTape(n_t, params.sim_days + params.buffer_days, params.sim_days, 512)
    with _engine(params, keep, "f32") as eng:
        split = eng.run_host(2 * n_t, seed=99)
        info = eng.launch_info()
        assert "prod" in info["kernel"] and "rules" in info["kernel"] and info["kernel_launches"] == 2
        odd = eng.run_host(7, seed=99)                                        # odd count: last twin computed, not stored
        off = eng.run_host(10, seed=99, first_trial=4)
        gen = eng.run_host(2 * n_t, seed=99, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
    monkeypatch.setenv("LETF_NO_SPLIT", "1")
    with _engine(params, keep, "f32") as eng:
        fused = eng.run_host(2 * n_t, seed=99)
        assert "prod" not in eng.launch_info()["kernel"]
    o = orc.run(params, rec)
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=2e-3)
    assert np.median(np.abs(split.final_wealth / fused.final_wealth - 1)) < 5e-6
    assert np.mean(split.num_trades != fused.num_trades) < 0.01
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(split, k), getattr(fused, k))
    np.testing.assert_array_equal(odd.final_wealth, split.final_wealth[:, :7])
    np.testing.assert_array_equal(off.final_wealth, split.final_wealth[:, 4:14])
    np.testing.assert_array_equal(off.num_trades, split.num_trades[:, 4:14])
    rel = np.abs(split.final_wealth - o.final_wealth) / np.abs(o.final_wealth)
    assert np.median(rel) < 1e-4 and np.mean(rel < 1e-2) > 0.97, (np.median(rel), np.mean(rel < 1e-2))
    assert np.mean(split.num_trades != o.num_trades) < 0.03
    np.testing.assert_array_equal(split.regime_days, o.regime_days)
    np.testing.assert_array_equal(split.start_offset, o.start_offset)
    np.testing.assert_array_equal(split.jump_days, o.jump_days)


def test_edge_sizes_and_offsets(engine_lib):
    """empty launch, single trial, odd trial counts and odd first_trial with antithetic pairing, the tail of the
    two-trials-per-thread fast path"""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(["S1", "S2", "S4"], 1, synthetic.fallback_regime_model("default"), ns)
    with _engine(params, keep, "f32") as eng:
        assert eng.run_host(0).final_wealth.shape == (3, 0)
        whole = eng.run_host(11, seed=5)
        one = eng.run_host(1, seed=5, first_trial=7)
        odd = eng.run_host(4, seed=5, first_trial=3)
    np.testing.assert_array_equal(one.final_wealth[:, 0], whole.final_wealth[:, 7])
    np.testing.assert_array_equal(odd.final_wealth, whole.final_wealth[:, 3:7])
    np.testing.assert_array_equal(odd.num_trades, whole.num_trades[:, 3:7])
    np.testing.assert_array_equal(whole.start_offset[0::2][:5], whole.start_offset[1::2])     # twins share sim_id
    # fast path (c2 shape): odd counts exercise the padding lane
    ns2 = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                     USE_ANTITHETIC_VARIATES=False)
    p2, k2 = build_params(["S1", "B_UPRO", "S3"], 1, synthetic.normal_regime_model("c2"), ns2)
    with _engine(p2, k2, "f32") as eng:
        a = eng.run_host(1001, seed=9)
        b = eng.run_host(1, seed=9, first_trial=1000)
        c = eng.run_host(64, seed=9, first_trial=500)
    np.testing.assert_array_equal(a.final_wealth[:, 1000], b.final_wealth[:, 0])
    np.testing.assert_array_equal(a.final_wealth[:, 500:564], c.final_wealth)
    np.testing.assert_array_equal(a.max_dd[:, 500:564], c.max_dd)
    assert (a.flags == 0).all() and (a.regime_days[0] == p2.sim_days).all()


def test_horizon_leverage_grid_config5(engine_lib):
    """BASELINE configs[4] shape at a small size: per-cell histogram quantiles (device reduction) agree with exact
    percentiles of the per-trial outputs, the vol-drag ordering holds, and the straight-line t(5)/t(10) loop ran."""
    from letf_monte_carlo_b200.grid import HorizonGrid
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                    USE_ANTITHETIC_VARIATES=False)
    sids = ["B_UPRO", "S3", "S2"]
    model = synthetic.normal_regime_model("c2")
    grid = HorizonGrid([1, 5, 10], sids, model, ns, precision="f32", device=0)
    n = 20000
    rows = grid.run(n, seed=4).table()
    assert len(rows) == 9 and all(r["valid"] == n for r in rows)
    assert grid.engines[0].launch_info()["regs_per_thread"] > 128          # the two-trials-per-thread build
    for h, eng in zip(grid.horizons, grid.engines):
        exact = eng.run_host(n, seed=4)
        for s, sid in enumerate(sids):
            row = next(r for r in rows if r["horizon"] == h and r["id"] == sid)
            for q in (10, 50, 90):
                want = np.percentile(exact.final_wealth[s], q)
                assert abs(np.log10(row[f"p{q}"] / want)) < 2 * 30 / 4096, (h, sid, q)
            assert abs(row["dd_median"] - np.median(exact.max_dd[s])) < 2 / 4096
    # mean log-growth per day against the closed form L mu - costs - (L sigma)^2 / 2 (second order in the drift)
    g = {(r["horizon"], r["id"]): r["mean_log10_growth"] * np.log(10) / (252 * r["horizon"]) for r in rows}
    vol, mu = 0.16 / np.sqrt(252), 0.08 / 252
    for sid, a in (("B_UPRO", "UPRO"), ("S3", "SSO"), ("S2", "SPY")):
        c = ns.ASSETS[a]; L = c["leverage"]
        m = L * mu - c["expense_ratio"] / 252 - (L - 1) * (0.045 + c["borrow_spread"]) / 252
        want = m - 0.5 * (L * vol) ** 2 - 0.5 * m * m
        se = L * vol / np.sqrt(2520 * n)
        assert abs(g[(10.0, sid)] - want) < 5 * se + 2e-6, (sid, g[(10.0, sid)], want)
    grid.close()


def test_bootstrap_lean_kernel_vs_generic_and_oracle(engine_lib):
    """legacy_hybrid block bootstrap: bench.py's `boot` workload runs the lean buy&hold build.
    Chain of evidence on the device: lean (Philox) == generic fp32 build recording the same counters (identical
    block choices and regime days, fp32 reassociation only on wealth), and the oracle replaying that tape agrees."""
    sids = ["S1", "S2", "S3"]
    doc = {"case": dict(variant="boot", years=3, n_tapes=128, seed=0, cfg={"SIM_ENGINE_MODE": "legacy_hybrid"},
                        strategies=sids)}
    ns, model, params, keep, _ = case_inputs(doc)
    assert params.market_mode == abi.MARKET_BOOTSTRAP
    n_t = 128
    n = 2 * n_t if params.antithetic_pairs else n_t
    G = params.sim_days + params.buffer_days
    from letf_monte_carlo_b200.tape import boot_tape_shape
    rec = HostTape(n_t, G, params.sim_days, 512, *boot_tape_shape(params))
    with _engine(params, keep, "f32") as eng:
        lean = eng.run_host(n, seed=2718)
        regs_lean = eng.launch_info()["regs_per_thread"]
        gen = eng.run_host(n, seed=2718, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
        regs_gen = eng.launch_info()["regs_per_thread"]
    assert regs_lean != regs_gen                     # two different instantiations really ran
    o = orc.run(params, rec)
    np.testing.assert_array_equal(lean.regime_days, gen.regime_days)
    np.testing.assert_array_equal(gen.regime_days, o.regime_days)
    np.testing.assert_array_equal(lean.start_offset, o.start_offset)
    np.testing.assert_allclose(lean.final_wealth, gen.final_wealth, rtol=5e-5)
    np.testing.assert_allclose(gen.final_wealth, o.final_wealth, rtol=3e-4)
    np.testing.assert_allclose(lean.max_dd, o.max_dd, atol=1e-4)
    # antithetic twins are identical paths outside the joint-t generator (calibration.py:563-564 is its only reader)
    np.testing.assert_array_equal(lean.final_wealth[:, 0::2], lean.final_wealth[:, 1::2])


def test_sampled_tax_path_replays_trade_lists_on_device(engine_lib):
    """parallel_monte_carlo_fixed -> MonteCarloRun.taxable_summary.  Only the <= 50 wealth-ranked
    trials per strategy are re-simulated for their Trade_List (reporting.py:457-467); each replayed list must carry
    exactly the journal entries the bulk launch counted for that trial."""
    from letf_monte_carlo_b200 import parallel_monte_carlo_fixed
    ns = synthetic.config_namespace(NUM_SIMULATIONS=160, GPU_PRECISION="f64")
    sids = ["S2", "S7", "S8", "S19"]
    run = parallel_monte_carlo_fixed(sids, 1, synthetic.fallback_regime_model(), {}, cfg=ns, verbose=False, device=0)
    bulk = run.engine.run_host(160, seed=run.seed, journal=True)
    seen = []

    def tax(trades, years, capital, strategy_id, tax_config):
        seen.append((strategy_id, len(trades)))
        return 0.2 * sum(t["dollar_amount"] for t in trades if t["action"] == "SELL") * 0.01
    rows = run.taxable_summary(tax_fn=tax, tax_config={}, cfg=ns)
    assert sorted(r["id"] for r in rows) == ["S19", "S7", "S8"]
    for r in rows:
        s = sids.index(r["id"])
        assert len(r["sampled_trials"]) == 50
        batch = run[r["id"]].trade_lists(r["sampled_trials"])          # one launch for the 50 trials
        assert [len(b) for b in batch] == [int(bulk.journal_count[s, i]) for i in r["sampled_trials"]]
        for i, b in list(zip(r["sampled_trials"], batch))[::7]:
            assert run[r["id"]].trade_list(i) == b                     # identical to the single-trial replay
        assert r["pcts"]["p60"]["post_wealth"] <= r["pcts"]["p60"]["pre_wealth"] + 1e-9
        assert 0.0 <= r["win"] <= 100.0
    assert len(seen) <= 150 and all(n > 0 for _, n in seen)


def test_typed_series_kernel_vs_generic_and_oracle(engine_lib):
    """materialised-path mode with the default 19-strategy list: the typed-slot fp32 series build (bench.py
    `series19`) against the generic fp32 series build (journal on -> tape-capable kernel) and the fp64 oracle."""
    import ctypes as C
    rng = np.random.default_rng(11)
    T, n = 900, 96
    ns = synthetic.config_namespace()
    params, keep = build_params(ALL19, T / 252, synthetic.fallback_regime_model(), ns)
    rets = {a: np.ascontiguousarray(rng.standard_t(5, (T, n)) * s + 0.0004) for a, s in
            zip(abi.ASSET_NAMES, (0.028, 0.028, 0.019, 0.022, 0.0095, 0.012))}
    vix = np.ascontiguousarray(14 + 12 * rng.random((T, n)))
    regime = np.ascontiguousarray((rng.random((T, n)) < 0.2).astype(np.int32))
    from oracle.oracle import HostOutputs
    o = HostOutputs(params.n_strategies, n)
    s = abi.Series()
    s.n_paths, s.sim_days = n, T
    for a, name in enumerate(abi.ASSET_NAMES):
        s.asset_ret[a] = rets[name].ctypes.data
    s.vix, s.regime = vix.ctypes.data, regime.ctypes.data
    assert orc.lib().letf_oracle_run_strategies(C.byref(params), C.byref(s), C.byref(o.as_struct())) == 0
    with _engine(params, keep, "f32") as eng:
        typed = eng.run_strategies(rets, vix=vix, regime=regime).to_host()
        regs_typed = eng.launch_info()["regs_per_thread"]
        gen = eng.run_strategies(rets, vix=vix, regime=regime, journal=True).to_host()
        regs_gen = eng.launch_info()["regs_per_thread"]
    assert regs_typed != regs_gen                    # two different instantiations really ran
    assert np.mean(typed.num_trades != gen.num_trades) < 0.01
    assert np.mean(gen.num_trades != o.num_trades) < 0.02
    same = typed.num_trades == o.num_trades          # a flipped trade decision moves the whole path: compare the rest
    assert np.median(np.abs(typed.final_wealth / o.final_wealth - 1)[same]) < 1e-5
    np.testing.assert_allclose(typed.final_wealth[same], o.final_wealth[same], rtol=5e-4)
    np.testing.assert_allclose(typed.final_wealth[:3], o.final_wealth[:3], rtol=2e-5)      # buy & hold rows


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_bootstrap_philox_distribution_matches_live_reference(precision, engine_lib):
    """KS + summary-quantile agreement of the on-device block bootstrap with the LIVE reference's own sampler and NumPy
    streams (tests/golden/ref_distribution_boot_1y.npz: 8 000 trials, legacy_hybrid, 1 year, 10 strategies)."""
    from tests.common import assert_same_distribution, bootstrap_distribution_case
    ref, sids, params, keep = bootstrap_distribution_case()
    with _engine(params, keep, precision) as eng:
        out = eng.run_host(40000, seed=1618)
    assert_same_distribution(ref, out, sids)


def test_historical_anchor_start_matches_live_reference(engine_lib):
    """RANDOM_START_METHOD = 'historical_anchor': Bernoulli start regime on the device against the reference's uniformly
    drawn history row (tests/golden/ref_distribution_anchor_1y.npz, 4 000 live-reference trials)."""
    from tests.common import anchor_distribution_case, assert_same_distribution
    ref, sids, params, keep = anchor_distribution_case()
    with _engine(params, keep, "f32") as eng:
        out = eng.run_host(40000, seed=577)
    assert_same_distribution(ref, out, sids)
    assert stats.ks_2samp(ref["high_vol_days"], out.regime_days[1]).pvalue > 1e-3


def test_public_api_bootstrap_via_historical_df(engine_lib):
    """mc_runner.py:79-84: parallel_monte_carlo_fixed(..., historical_df) builds the sampler itself; in legacy_hybrid mode
    that selects the block-bootstrap generator.  Same results as handing the sampler over in regime_model; the engine
    handle is reused across calls with the same tables (content-keyed cache)."""
    from letf_monte_carlo_b200 import parallel_monte_carlo_fixed
    from letf_monte_carlo_b200.bootstrap import create_bootstrap_sampler
    hist = synthetic.synthetic_history(n_days=8000, seed=9)
    ns = synthetic.config_namespace(NUM_SIMULATIONS=300, SIM_ENGINE_MODE="legacy_hybrid")
    sids = ["S1", "S2", "S4", "S9"]
    a = parallel_monte_carlo_fixed(sids, 2, synthetic.fallback_regime_model(), {}, hist, cfg=ns, verbose=False, device=0)
    assert a.engine.params.market_mode == abi.MARKET_BOOTSTRAP and a.engine.params.boot.hist_len == 8000
    model = synthetic.fallback_regime_model()
    model["bootstrap_sampler"] = create_bootstrap_sampler(hist, cfg=ns)
    b = parallel_monte_carlo_fixed(sids, 2, model, {}, cfg=ns, verbose=False, device=0)
    assert b.engine is a.engine
    for sid in sids:
        np.testing.assert_array_equal(a[sid].final_wealth, b[sid].final_wealth)
        assert np.isfinite(a[sid].final_wealth).all() and (a[sid].final_wealth > 0).all()
    assert a["S4"].num_trades.sum() > 0
    inst = parallel_monte_carlo_fixed(sids, 2, synthetic.fallback_regime_model(), {}, hist,
                                      cfg=synthetic.config_namespace(NUM_SIMULATIONS=300), verbose=False, device=0)
    assert inst.engine.params.market_mode == abi.MARKET_JOINT_T      # institutional_v1 + joint model wins (engine.py:45)


def test_run_strategy_fixed_dropin_matches_live_reference(engine_lib):
    """letf_monte_carlo_b200.run_strategy_fixed against the LIVE reference's run_strategy_fixed (strategy.py:55) on one
    seeded frame (tests/golden/run_strategy_fixed_700d.npz): daily equity curve, trade count, journal entries."""
    import os
    from letf_monte_carlo_b200 import TradeJournal, run_strategy_fixed
    from letf_monte_carlo_b200.trade import TAXABLE_IDS
    from oracle.make_strategy_golden import strategy_frame
    from tests.common import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "run_strategy_fixed_700d.npz"))
    df, regime = strategy_frame(int(g["T"]), int(g["seed"]))
    for sid in [str(x) for x in g["sids"]]:
        j = TradeJournal() if sid in TAXABLE_IDS else None
        ec, nt = run_strategy_fixed(df, sid, regime, {}, apply_costs=True, trade_journal=j, device=0)
        ref = g[f"{sid}_curve"]
        assert len(ec) == len(ref) and list(ec.index) == list(df.index)
        np.testing.assert_allclose(ec.values, ref, rtol=1e-11, err_msg=sid)
        assert nt == int(g[f"{sid}_trades"]), sid
        if j is not None:
            tl = j.get_full_trades()
            assert len(tl) == int(g[f"{sid}_journal_n"]), sid
            np.testing.assert_array_equal([t["day_index"] for t in tl], g[f"{sid}_journal_days"])
            assert abs(sum(t["dollar_amount"] for t in tl) - float(g[f"{sid}_journal_volume"])) <= 1e-9 * max(1.0, float(g[f"{sid}_journal_volume"]))


def test_simulate_single_path_fixed_dropin_and_equity_curve(engine_lib):
    """the single-trial seam (engine.py:423) returns the bulk launch's numbers for that trial, with the reference's keys;
    the replayed equity curve ends at Final_Wealth and reproduces Max_DD (engine.py:578-586)."""
    from letf_monte_carlo_b200 import parallel_monte_carlo_fixed, simulate_single_path_fixed
    from letf_monte_carlo_b200.mc_runner import _engine_for
    ns = synthetic.config_namespace(NUM_SIMULATIONS=64, GPU_PRECISION="f64")
    sids = ["S1", "S2", "S4", "S8", "S19"]
    model = synthetic.fallback_regime_model()
    bulk = parallel_monte_carlo_fixed(sids, 1, model, {}, cfg=ns, verbose=False, device=0)
    for sim_id, anti in ((6, False), (6, True), (20, True)):
        one = simulate_single_path_fixed((sim_id, 1, model, {}, sids, anti), cfg=ns, device=0)
        k = sim_id | (1 if anti else 0)
        for sid in sids:
            r = one[sid]
            assert r["Final_Wealth"] == bulk[sid][k]["Final_Wealth"] and r["Num_Trades"] == bulk[sid][k]["Num_Trades"]
            assert r["Max_DD"] == bulk[sid][k]["Max_DD"] and r["Metadata"]["regime_counts"] == bulk[sid][k]["Metadata"]["regime_counts"]
            assert ("Trade_List" in r) == (sid in ("S8", "S19"))
    eng = bulk.engine
    for s, sid in enumerate(sids):
        ec = eng.equity_curve(7, s, seed=bulk.seed)
        assert len(ec) == 252 and ec[-1] == pytest.approx(bulk[sid][7]["Final_Wealth"], rel=1e-13)
        rm = np.maximum.accumulate(ec)
        assert abs(((ec - rm) / np.where(rm > 0, rm, 1.0)).min()) == pytest.approx(bulk[sid][7]["Max_DD"], abs=1e-12)


def test_summary_only_route_matches_per_trial_route(engine_lib):
    """parallel_monte_carlo_fixed(..., summary_only=True): reduction on the device, only the packed summary leaves it.  Its
    table must agree with the reference's definitions (reporting.py:353-388) evaluated on the per-trial arrays of the same
    run: histogram percentiles to bin width, exact=True percentiles bit for bit; device_results=True copies lazily."""
    from letf_monte_carlo_b200 import parallel_monte_carlo_fixed
    from letf_monte_carlo_b200.summary import summary_table
    ns = synthetic.config_namespace(NUM_SIMULATIONS=6000, GPU_SEED=7)
    sids = ["S1", "S2", "S4", "S19"]
    model = synthetic.fallback_regime_model("default")
    full = parallel_monte_carlo_fixed(sids, 2, model, {}, cfg=ns, verbose=False)
    ref = {r["id"]: r for r in summary_table(full, 2)}
    sm = parallel_monte_carlo_fixed(sids, 2, model, {}, cfg=ns, verbose=False, summary_only=True)
    assert sm.d2h_bytes < 1 << 20
    approx = {r["id"]: r for r in sm.table()}
    exact = {r["id"]: r for r in sm.table(exact=True)}
    for sid in sids:
        for k in ("p10", "p25", "p40", "p60", "p75", "p90", "median"):
            assert exact[sid][k] == ref[sid][k], (sid, k)
            assert abs(np.log10(approx[sid][k] / ref[sid][k])) < 2 * 30 / 4096, (sid, k)
        assert exact[sid]["dd"] == ref[sid]["dd"] and exact[sid]["cagr"] == ref[sid]["cagr"]
        assert abs(approx[sid]["win"] - ref[sid]["win"]) < 1e-9 and abs(approx[sid]["trades"] - ref[sid]["trades"]) < 1e-9
        assert approx[sid]["valid"] == 6000
    lazy = parallel_monte_carlo_fixed(sids, 2, model, {}, cfg=ns, verbose=False, device_results=True)
    assert lazy._out is None
    np.testing.assert_array_equal(lazy["S4"].final_wealth, full["S4"].final_wealth)
    assert lazy["S19"][5]["Num_Trades"] == full["S19"][5]["Num_Trades"]


def test_reduce_on_a_second_engine_after_the_first(engine_lib):
    """the reduction's shared-memory opt-in is set per launch (it is a per-device attribute): two engines in one process,
    each with its own summary, both reduce"""
    ns = synthetic.config_namespace()
    params, keep = build_params(["S1", "S2"], 1, synthetic.fallback_regime_model("default"), ns)
    for _ in range(2):
        with _engine(params, keep, "f32") as eng:
            out = eng.run(512, seed=3)
            sm = eng.new_summary(bench_strategy=1)
            eng.reduce(out, sm)
            assert int(sm.counters[0, 0]) == 512 and int(sm.hist[0, 0].sum()) == 512
            info = eng.launch_info()
            assert info["kernel_launches"] == 1 and info["aux_launches"] == 1 and info["kernel"]


def test_run_strategy_fixed_infers_the_regime_path_like_the_reference(engine_lib):
    """run_strategy_fixed(df, sid, None, ...) and a regime path of the wrong length: the reference infers the path from VIX /
    realised volatility / term spread (strategy.py:64-78) - cash rate and spread cost follow it.  Equity curves against the
    live reference (tests/golden/infer_regime_900d.npz)."""
    import os
    from letf_monte_carlo_b200 import run_strategy_fixed
    from oracle.make_regime_inference_golden import frame
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "infer_regime_900d.npz"))
    df = frame()
    df["VIX"] = df["VIX"].ffill()
    for sid in ("S4", "S8", "S16"):
        curve, nt = run_strategy_fixed(df, sid, None, {}, apply_costs=True)
        np.testing.assert_allclose(curve.to_numpy(), g[f"{sid}_curve"], rtol=1e-11)
        assert nt == int(g[f"{sid}_trades"])
        curve2, _ = run_strategy_fixed(df, sid, np.zeros(17, dtype=int), {}, apply_costs=True)
        np.testing.assert_allclose(curve2.to_numpy(), g[f"{sid}_curve_mismatch"], rtol=1e-11)
    # the engine-level seam refuses columns whose length is not the frame's (it would read past them)
    from letf_monte_carlo_b200.params import build_params as bp
    params, keep = bp(["S4"], len(df) / 252.0, {"regime_params": {}, "transition_matrix": np.eye(2), "joint_return_model": None},
                      synthetic.config_namespace())
    with _engine(params, keep, "f64") as eng:
        with pytest.raises(ValueError):
            eng.run_strategy_curve({"SPY": df["SPY_Ret"].to_numpy(), "TQQQ": df["TQQQ_Ret"].to_numpy()}, 0, regime=np.zeros(17, np.int32))
        with pytest.raises(ValueError):
            eng.run_strategy_curve({"SPY": df["SPY_Ret"].to_numpy(), "TQQQ": df["TQQQ_Ret"].to_numpy()}, 0,
                                   regime=np.full(len(df), 2, np.int32))


@pytest.mark.parametrize("anti", [False, True])
def test_simulate_joint_returns_t_dropin_matches_live_reference(anti, engine_lib):
    """the last inner seam (calibration.py:480-590): per-day SPY / QQQ / TLT returns of one path along a caller-supplied regime
    path, variates injected through `rng`; fp64 device kernel vs the live reference on the same variates, <= 1e-12"""
    import os
    from letf_monte_carlo_b200 import simulate_joint_returns_t
    from oracle.make_joint_returns_golden import inputs
    from tests.test_kernel_emulation import _PresetRng
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "joint_returns_400d.npz"))
    regime, x, chi2, joint = inputs()
    for mm in (True, False):
        ns = synthetic.config_namespace(USE_MOMENT_MATCHING=mm)
        out = simulate_joint_returns_t(len(regime), regime, joint, _PresetRng(x, chi2), antithetic=anti, cfg=ns)
        assert list(out) == ["SPY_Ret", "QQQ_Ret", "TLT_Ret"]
        got = np.stack([out[a] for a in out], axis=1)
        np.testing.assert_allclose(got, g[f"ret_mm{int(mm)}_anti{int(anti)}"], rtol=1e-12, atol=1e-15)
    # with a real generator: finite, right length, reproducible for a seeded stream
    a = simulate_joint_returns_t(300, regime[:300], joint, np.random.default_rng(5), cfg=synthetic.config_namespace())
    b = simulate_joint_returns_t(300, regime[:300], joint, np.random.default_rng(5), cfg=synthetic.config_namespace())
    assert all(len(v) == 300 and np.isfinite(v).all() for v in a.values())
    np.testing.assert_array_equal(a["QQQ_Ret"], b["QQQ_Ret"])


@pytest.mark.parametrize("days,random_start", [(2, True), (5, False), (61, True), (257, False), (300, True)])
def test_pair_and_split_paths_at_edge_horizons(days, random_start, engine_lib, monkeypatch):
    """horizons around the staging depth of the rule kernel (3 days ahead), the rolling-window lengths and the steady-state
    switch of its day body (256), with and without the random-start buffer: the split pipeline and the pair kernel against
    the fused / general builds on the same Philox counters"""
    model = synthetic.fallback_regime_model("default")
    years = days / 252.0 + 1e-9
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=random_start)
    p3, k3 = build_params(["S2", "B_QQQ", "B_TLT"], years, model, ns)
    p19, k19 = build_params(ALL19, years, model, synthetic.config_namespace(USE_RANDOM_START=random_start))
    assert p3.sim_days == days
    with _engine(p3, k3, "f32") as eng:
        pair = eng.run_host(300, seed=3)
        assert eng.launch_info()["kernel"].startswith("f32_joint_t_pair")
    with _engine(p19, k19, "f32") as eng:
        split = eng.run_host(300, seed=4)
        assert "rules" in eng.launch_info()["kernel"]
    monkeypatch.setenv("LETF_NO_JPAIR", "1")
    monkeypatch.setenv("LETF_NO_SPLIT", "1")
    with _engine(p3, k3, "f32") as eng:
        lean = eng.run_host(300, seed=3)
    with _engine(p19, k19, "f32") as eng:
        fused = eng.run_host(300, seed=4)
    np.testing.assert_allclose(pair.final_wealth, lean.final_wealth, rtol=5e-5)
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=2e-3)
    for a, b in ((pair, lean), (split, fused)):
        for k in ("flags", "start_offset", "regime_days", "jump_days"):
            np.testing.assert_array_equal(getattr(a, k), getattr(b, k))
    assert np.mean(split.num_trades != fused.num_trades) < 0.01


@pytest.mark.parametrize("sids,kname", [(["S2", "B_QQQ", "B_TLT"], "f32_joint_t_pair_c3"), (["S1", "S2"], "f32_joint_t_pair_2tedf"),
                                        (ALL19, "prod_sweep19_tedf")], ids=["calibrated_nu", "calibrated_te_pair", "calibrated_te_split"])
def test_pair_kernels_with_calibrated_models(sids, kname, engine_lib, monkeypatch):
    """calibrated models (what letf.run() feeds the path): non-integer joint dof (calibration.py:399; Marsaglia-Tsang chi-square
    with out-of-line retries in the pair kernel) and a tracking_residual_model with non-integer dof and a tanh cap (run-time
    sampler builds): specialised kernels against the general builds on the same Philox counters, and against the oracle"""
    model = synthetic.fallback_regime_model("vix_te" if "tedf" in kname else "default")
    model["joint_return_model"]["regimes"][0]["nu"] = 5.3
    model["joint_return_model"]["regimes"][1]["nu"] = 4.2
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, 1, model, ns)
    n_t = 64
    rec = HostTape(n_t, params.sim_days + params.buffer_days, params.sim_days, 512)
    with _engine(params, keep, "f32") as eng:
        fast = eng.run_host(2 * n_t, seed=123)
        assert kname in eng.launch_info()["kernel"], eng.launch_info()["kernel"]
        eng.run_host(2 * n_t, seed=123, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
    monkeypatch.setenv("LETF_NO_JPAIR", "1")
    monkeypatch.setenv("LETF_NO_SPLIT", "1")
    with _engine(params, keep, "f32") as eng:
        slow = eng.run_host(2 * n_t, seed=123)
    o = orc.run(params, rec)
    np.testing.assert_allclose(fast.final_wealth, slow.final_wealth, rtol=2e-3)
    assert np.median(np.abs(fast.final_wealth / slow.final_wealth - 1)) < 5e-6
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(fast, k), getattr(slow, k))
    rel = np.abs(fast.final_wealth - o.final_wealth) / np.abs(o.final_wealth)
    assert np.median(rel) < 1e-4 and np.mean(rel < 1e-2) > 0.97
    np.testing.assert_array_equal(fast.regime_days, o.regime_days)
    np.testing.assert_array_equal(fast.jump_days, o.jump_days)


@pytest.mark.parametrize("sids,model,kname", [
    (["S1", "S4", "S7", "S10", "S15"], "default", "prod_sweep19+rules_sub19"),
    (["S19", "S2", "S8", "S6"], "default", "prod_sweep19+rules_sub19"),
    (["S3", "S1", "S2", "B_TMF", "B_TLT", "S12"], "default", "prod_sweep19+rules_sub19"),
    (["S5", "S9", "S17"], "vix_te", "prod_sweep19_tedf+rules_sub19"),
    (["S1", "S3", "S2", "B_TMF"], "default", "prod_sweep19"),
    (["S4", "S2"], "default", "prod_c4+rules_sub6"),
    (["S19", "S5", "B_TMF", "S1", "S2", "B_TLT"], "default", "prod_c4+rules_sub6"),
    (["S6", "S4"], "vix_te", "prod_c4_tedf+rules_sub6"),
    (["S2", "S3", "B_QQQ", "B_UPRO", "B_TLT", "S9", "S13"], "default", "prod_all+rules_all8"),
    (["S2", "S3", "B_QQQ", "B_UPRO", "B_TLT", "S9", "S13", "S4", "S5", "S7"], "default", "prod_all+rules_all"),
    (["S1", "S3", "S2", "B_UPRO", "B_TMF"], "default", "prod_all"),
    (["S4", "S6", "S12", "S18", "B_QQQ"], "vix_te", "prod_all_tedf+rules_all8"),
], ids=["routed_subset", "routed_out_of_order", "routed_extra_buy_hold", "routed_calibrated_te", "routed_buy_hold_only", "routed6_sma_vs_hold",
        "routed6_out_of_order", "routed6_calibrated_te", "six_series",
        "six_series_ten_slots",
        "buy_hold_five_series", "generic_calibrated_te"])
def test_generic_split_pipeline_vs_fused_and_oracle(sids, model, kname, engine_lib, monkeypatch):
    """strategy lists without a typed build.  A subset of the default sweep runs the 19-rule kernel with the other slots
    switched off and its outputs routed (slot list 3); anything else the generic split pipeline (uncompacted pair producer ->
    six-field records -> dynamic-slot rule kernel; producer alone for a buy&hold-only list).  Both against the fused generic
    build on the same Philox counters and against the oracle on the recorded tape."""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, 2, synthetic.fallback_regime_model(model), ns)
    n_t = 160
    rec = HostTape(n_t, params.sim_days + params.buffer_days, params.sim_days, 512)
    with _engine(params, keep, "f32") as eng:
        fast = eng.run_host(2 * n_t, seed=77)
        info = eng.launch_info()
        assert info["kernel"].endswith(kname), info["kernel"]
        eng.run_host(2 * n_t, seed=77, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
        odd = eng.run_host(2 * n_t - 3, seed=77)
    monkeypatch.setenv("LETF_NO_GENERIC_SPLIT", "1")
    with _engine(params, keep, "f32") as eng:
        slow = eng.run_host(2 * n_t, seed=77)
        assert "prod_" not in eng.launch_info()["kernel"]
    if "rules_sub" in kname:    # the generic pipeline serves the same request when the routed one is switched off
        monkeypatch.delenv("LETF_NO_GENERIC_SPLIT")
        monkeypatch.setenv("LETF_NO_ROUTED_SPLIT", "1")
        with _engine(params, keep, "f32") as eng:
            gen = eng.run_host(2 * n_t, seed=77)
            assert "rules_all8" in eng.launch_info()["kernel"]
        np.testing.assert_allclose(fast.final_wealth, gen.final_wealth, rtol=2e-3)
        np.testing.assert_array_equal(fast.flags, gen.flags)
    o = orc.run(params, rec)
    np.testing.assert_allclose(fast.final_wealth, slow.final_wealth, rtol=2e-3)
    assert np.median(np.abs(fast.final_wealth / slow.final_wealth - 1)) < 5e-6
    np.testing.assert_allclose(fast.max_dd, slow.max_dd, atol=1e-4)
    assert np.mean(fast.num_trades != slow.num_trades) < 0.02
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(fast, k), getattr(slow, k))
    np.testing.assert_array_equal(odd.final_wealth, fast.final_wealth[:, :2 * n_t - 3])
    rel = np.abs(fast.final_wealth - o.final_wealth) / np.abs(o.final_wealth)
    assert np.median(rel) < 1e-4 and np.mean(rel < 1e-2) > 0.97
    np.testing.assert_array_equal(fast.regime_days, o.regime_days)
    np.testing.assert_array_equal(fast.jump_days, o.jump_days)


@pytest.mark.parametrize("sids,kname", [(ALL19, "series_records+rules_sweep19"),
                                        (["S19", "S2", "S8", "S6", "B_TLT"], "series_records+rules_sub19"),
                                        (["S4", "S9", "B_QQQ", "B_UPRO", "S2"], "series_records+rules_all8"),
                                        (["S4", "S9", "B_QQQ", "B_UPRO", "S2", "S5", "S7", "S11", "S12"], "series_records+rules_all"),
                                        (["S1", "S2", "S3"], "series_records")],
                         ids=["typed", "routed", "dynamic", "dynamic_nine_slots", "buy_hold_only"])
def test_series_mode_through_the_record_pass(sids, kname, engine_lib, monkeypatch):
    """materialised-path mode (run_strategy_fixed seam) through the split pipeline: the bandwidth-bound record pass over the
    caller's fp64 series + the rule kernel of the strategy list, against the fused series builds (LETF_NO_SPLIT=1) and the
    fp64 oracle on the same frame; buy&hold rows, flags and regime counts are bit-identical between the two routes"""
    import ctypes as C
    rng = np.random.default_rng(17)
    T, n = 900, 333                      # (not a multiple of the block size)
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, T / 252, synthetic.fallback_regime_model(), ns)
    rets = {a: np.ascontiguousarray(rng.standard_t(5, (T, n)) * s + 0.0004) for a, s in
            zip(abi.ASSET_NAMES, (0.028, 0.028, 0.019, 0.022, 0.0095, 0.012))}
    vix = np.ascontiguousarray(14 + 12 * rng.random((T, n)))
    vix[100, 5] = 140.0
    regime = np.ascontiguousarray((rng.random((T, n)) < 0.2).astype(np.int32))
    tlt = np.ascontiguousarray(rng.normal(0, 0.008, (T, n)))
    from oracle.oracle import HostOutputs
    o = HostOutputs(params.n_strategies, n)
    s = abi.Series()
    s.n_paths, s.sim_days = n, T
    for a, name in enumerate(abi.ASSET_NAMES):
        s.asset_ret[a] = rets[name].ctypes.data
    s.tlt_ret, s.vix, s.regime = tlt.ctypes.data, vix.ctypes.data, regime.ctypes.data
    assert orc.lib().letf_oracle_run_strategies(C.byref(params), C.byref(s), C.byref(o.as_struct())) == 0
    with _engine(params, keep, "f32") as eng:
        split = eng.run_strategies(rets, tlt_ret=tlt, vix=vix, regime=regime).to_host()
        assert eng.launch_info()["kernel"] == kname, eng.launch_info()["kernel"]
        # a frame with its own Cash_Ret column (what the reference's frames carry, strategy.py:138): copied beside the records
        cash = np.ascontiguousarray(2e-4 + 1e-4 * rng.random((T, n)))
        split_c = eng.run_strategies(rets, tlt_ret=tlt, vix=vix, regime=regime, cash_ret=cash).to_host()
        assert eng.launch_info()["kernel"] == kname.replace("rules_sweep19", "rules_sub19")     # (the typed kernels read no cash column)
    monkeypatch.setenv("LETF_NO_SPLIT", "1")
    with _engine(params, keep, "f32") as eng:
        fused = eng.run_strategies(rets, tlt_ret=tlt, vix=vix, regime=regime).to_host()
        assert "series_records" not in eng.launch_info()["kernel"]
        fused_c = eng.run_strategies(rets, tlt_ret=tlt, vix=vix, regime=regime, cash_ret=cash).to_host()
    np.testing.assert_allclose(split_c.final_wealth, fused_c.final_wealth, rtol=5e-4)
    assert np.mean(split_c.num_trades != fused_c.num_trades) < 0.02
    hold = [k for k, sid in enumerate(sids) if ns.STRATEGIES[sid]["type"].startswith("benchmark")]
    np.testing.assert_array_equal(split.final_wealth[hold], fused.final_wealth[hold])
    np.testing.assert_array_equal(split.max_dd[hold], fused.max_dd[hold])
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(split, k), getattr(fused, k))
    assert split.flags[5] & abi.FLAG_REPAIRED
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=5e-4)
    assert np.mean(split.num_trades != fused.num_trades) < 0.02
    same = split.num_trades == o.num_trades
    assert np.mean(same) > 0.97
    assert np.median(np.abs(split.final_wealth / o.final_wealth - 1)[same]) < 1e-5
    rel = np.abs(split.final_wealth / o.final_wealth - 1)[same]      # (same trade count, a decision a day apart: a handful of cells)
    assert np.mean(rel < 5e-4) > 0.995 and rel.max() < 3e-2


def test_pair_kernel_and_split_pipeline_without_moment_matching(engine_lib, monkeypatch):
    """USE_MOMENT_MATCHING = False: the specialised paths still serve the shape (two passes, zero shift); against the fused
    one-pass builds on the same Philox counters and the oracle on the recorded tape"""
    model = synthetic.fallback_regime_model("default")
    for sids, kname in ((["S2", "B_QQQ", "B_TLT"], "f32_joint_t_pair_c3"), (ALL19, "prod_sweep19+rules_sweep19")):
        ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_MOMENT_MATCHING=False)
        params, keep = build_params(sids, 2, model, ns)
        n_t = 64
        rec = HostTape(n_t, params.sim_days + params.buffer_days, params.sim_days, 512)
        with _engine(params, keep, "f32") as eng:
            fast = eng.run_host(2 * n_t, seed=31)
            assert eng.launch_info()["kernel"].endswith(kname), eng.launch_info()["kernel"]
            eng.run_host(2 * n_t, seed=31, rng_mode=abi.RNG_PHILOX_RECORD, tape=rec)
        monkeypatch.setenv("LETF_NO_JPAIR", "1")
        monkeypatch.setenv("LETF_NO_SPLIT", "1")
        with _engine(params, keep, "f32") as eng:
            slow = eng.run_host(2 * n_t, seed=31)
        monkeypatch.delenv("LETF_NO_JPAIR")
        monkeypatch.delenv("LETF_NO_SPLIT")
        o = orc.run(params, rec)
        np.testing.assert_allclose(fast.final_wealth, slow.final_wealth, rtol=2e-3)
        assert np.median(np.abs(fast.final_wealth / slow.final_wealth - 1)) < 5e-6
        for k in ("flags", "start_offset", "regime_days", "jump_days"):
            np.testing.assert_array_equal(getattr(fast, k), getattr(slow, k))
        rel = np.abs(fast.final_wealth - o.final_wealth) / np.abs(o.final_wealth)
        assert np.median(rel) < 1e-4 and np.mean(rel < 1e-2) > 0.97
        np.testing.assert_array_equal(fast.regime_days, o.regime_days)


def test_split_launches_are_chunk_invariant(engine_lib, monkeypatch):
    """the split pipelines process trials in chunks of whole producer waves (memory-bounded); results must not depend on
    where the seams fall - one wave per chunk / 256 paths per chunk against the default chunking, bit for bit"""
    ns = synthetic.config_namespace()
    params, keep = build_params(ALL19, 30 / 252.0 + 1e-9, synthetic.fallback_regime_model("default"), ns)
    n = 250_001                                   # > 2 producer waves of 113 664 trials, odd
    with _engine(params, keep, "f32") as eng:
        whole = eng.run_host(n, seed=9)
        assert eng.launch_info()["kernel_launches"] == 2
    monkeypatch.setenv("LETF_SPLIT_WAVES", "1")
    with _engine(params, keep, "f32") as eng:
        parts = eng.run_host(n, seed=9)
        assert eng.launch_info()["kernel_launches"] >= 6
    for k in ("final_wealth", "max_dd", "num_trades", "flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(parts, k), getattr(whole, k))
    # materialised-path mode: 256 paths per chunk
    rng = np.random.default_rng(3)
    T, m = 120, 1000
    p2, k2 = build_params(["S19", "S2", "S8", "S6", "B_TLT"], T / 252, synthetic.fallback_regime_model(),
                          synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS))
    rets = {a: np.ascontiguousarray(rng.standard_t(5, (T, m)) * s) for a, s in zip(abi.ASSET_NAMES, (0.028, 0.028, 0.019, 0.022, 0.0095, 0.012))}
    vix = np.ascontiguousarray(14 + 12 * rng.random((T, m)))
    regime = np.ascontiguousarray((rng.random((T, m)) < 0.2).astype(np.int32))
    tlt = np.ascontiguousarray(rng.normal(0, 0.008, (T, m)))
    cash = np.ascontiguousarray(2e-4 + 1e-4 * rng.random((T, m)))
    monkeypatch.delenv("LETF_SPLIT_WAVES")
    with _engine(p2, k2, "f32") as eng:
        a = eng.run_strategies(rets, tlt_ret=tlt, vix=vix, regime=regime, cash_ret=cash).to_host()
        assert eng.launch_info()["kernel_launches"] == 2
    monkeypatch.setenv("LETF_SERIES_CHUNK", "256")
    with _engine(p2, k2, "f32") as eng:
        b = eng.run_strategies(rets, tlt_ret=tlt, vix=vix, regime=regime, cash_ret=cash).to_host()
        assert eng.launch_info()["kernel_launches"] == 8
    for k in ("final_wealth", "max_dd", "num_trades", "flags", "regime_days"):
        np.testing.assert_array_equal(getattr(b, k), getattr(a, k))


def test_trade_lists_of_several_strategies_in_one_launch(engine_lib):
    """the sampled-trial tax table replays the <= 50 trials of EVERY taxable strategy in one launch
    (letf_trade_lists_multi_host): same journals as the per-strategy batch and as the single-trial replay"""
    from letf_monte_carlo_b200 import parallel_monte_carlo_fixed
    ns = synthetic.config_namespace()
    ns.NUM_SIMULATIONS = 120
    sids = ["S2", "S4", "S7", "S9", "S19"]
    run = parallel_monte_carlo_fixed(sids, 2, synthetic.fallback_regime_model("default"), {}, cfg=ns, verbose=False)
    wanted = {"S7": [3, 50, 119], "S19": [0, 7], "S4": [5], "S9": []}          # S4 is not a taxable id: empty journal
    before = run.engine.launches
    multi = run.trade_lists_multi(wanted)
    assert run.engine.launches == before + 1
    assert [len(multi[s]) for s in wanted] == [3, 2, 1, 0]
    assert len(multi["S4"][0]) == 0
    for sid in ("S7", "S19"):
        batch = run[sid].trade_lists(wanted[sid])
        for a, b, i in zip(multi[sid], batch, wanted[sid]):
            assert len(a) > 0 and a == b and a == run[sid].trade_list(i)
    rows = run.taxable_summary()
    assert {r["id"] for r in rows} == {"S7", "S9", "S19"} and all(np.isfinite(r["sampled_tax"]).all() for r in rows)
