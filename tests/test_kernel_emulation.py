"""The fused kernel's per-trial body (csrc/letf_core.cuh), compiled for the HOST by tests/emu
(test-only build), against the oracle and the golden vectors.  This exercises the kernel
*source* on the CPU box; the same comparisons run on the device in test_gpu_parity.py."""
import numpy as np
import pytest
from scipy import stats

from letf_monte_carlo_b200 import abi, synthetic
from letf_monte_carlo_b200.params import build_params
from letf_monte_carlo_b200.tape import HostTape, make_tape
from oracle import oracle as orc
from tests.common import GOLDEN_NAMES, assert_matches_reference, case_inputs, load_case
from tests.emu import emu

ALL19 = ["S%d" % i for i in range(1, 20)]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_f64_kernel_body_reproduces_reference(name):
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    out = emu.run(params, tape, precision=abi.PREC_F64, rng_mode=abi.RNG_TAPE)
    # streaming fp64 vs the reference's array arithmetic: <= 1e-6 is the north-star bound, we hold 1e-11
    assert_matches_reference(out, doc, rtol=1e-11, dd_atol=1e-12)


def drifty_case(years=3, n=12, seed=7):
    """iid-normal market with mu = 1.2 sigma: runs of >= 20 non-negative returns are common, so the
    downside-vol windows hit pandas' recompute-on-collapse corner (oracle: orc_roll_std)."""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                    USE_ANTITHETIC_VARIATES=False)
    model = synthetic.normal_regime_model("c2")
    vol = np.array([0.16, 0.24, 0.12]) / np.sqrt(252)
    model["synthetic_normal"]["mu"] = list(1.2 * vol)
    params, keep = build_params(ALL19, years, model, ns)
    return params, keep, make_tape(params, n, seed)


def test_f64_constant_window_corner_matches_oracle():
    params, keep, tape = drifty_case()
    o = orc.run(params, tape)
    e = emu.run(params, tape, precision=abi.PREC_F64)
    np.testing.assert_allclose(e.final_wealth, o.final_wealth, rtol=1e-12)
    np.testing.assert_array_equal(e.num_trades, o.num_trades)
    np.testing.assert_array_equal(e.journal_count, o.journal_count)
    assert o.num_trades[9].min() > 0      # S10 (downside vol) traded


@pytest.mark.parametrize("name", ["inst_1y", "layers_c2_5y", "param_t_1y", "vix_te_1y", "voldrag_c1", "boot_1y"])
def test_f32_kernel_body_close_to_reference(name):
    """throughput mode (fp32 shocks / returns / signals, fp64 wealth compounding) on the SAME tape:
    stated tolerance 2e-4 relative on terminal wealth over <= 5 years, 1e-4 absolute on drawdown."""
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    out = emu.run(params, tape, precision=abi.PREC_F32, rng_mode=abi.RNG_TAPE)
    assert_matches_reference(out, doc, rtol=2e-4, dd_atol=1e-4, exact_counts=False)
    fw, dd, nt, jc, jv = __import__("tests.common", fromlist=["x"]).expected_arrays(doc)
    assert np.mean(out.num_trades != nt) < 0.02


@pytest.mark.parametrize("name", ["inst_30y", "layers_c2_30y"])
def test_f32_kernel_body_close_to_reference_at_30_years(name):
    """the benchmarked horizon (7 560 days): stated fp32 tolerance 5e-4 relative on terminal wealth (measured <= 7e-5)"""
    doc = load_case(name)
    ns, model, params, keep, tape = case_inputs(doc)
    out = emu.run(params, tape, precision=abi.PREC_F32, rng_mode=abi.RNG_TAPE)
    assert_matches_reference(out, doc, rtol=5e-4, dd_atol=5e-5, exact_counts=False)
    nt = __import__("tests.common", fromlist=["x"]).expected_arrays(doc)[2]
    assert np.sum(out.num_trades != nt) <= 3 and np.abs(out.num_trades - nt).max() <= 2


def _blank_tape(params, n_tapes, max_spells):
    from letf_monte_carlo_b200.tape import boot_tape_shape
    G = params.sim_days + params.buffer_days
    return HostTape(n_tapes, G, params.sim_days, max_spells, *boot_tape_shape(params))


@pytest.mark.parametrize("variant,years", [("default", 1), ("vix_te", 1), ("no_joint", 1), ("normal_c2", 2), ("boot", 1),
                                           ("regime_only", 1), ("boot_early", 1), ("boot_calm", 1)])
def test_philox_record_replays_through_oracle(variant, years):
    """RNG_PHILOX_RECORD writes every variate it consumes; feeding that tape to the oracle must give
    the same results -> the Philox path consumes draws exactly where the tape semantics say."""
    cfg_over = {"SIM_ENGINE_MODE": "legacy_hybrid"} if variant.startswith("boot") else {}
    if variant == "regime_only":      # random_start.py:169-172: the start regime is drawn, no buffer / offset
        variant, cfg_over = "default", {"RANDOM_START_METHOD": "regime_only", "START_REGIME_PROBABILITIES": {0: 0.5, 1: 0.5}}
    case = dict(variant=variant, years=years, n_tapes=3, seed=0, cfg=cfg_over,
                strategies=ALL19 if variant != "normal_c2" else ["S1", "S3", "B_UPRO", "S4", "S6", "S19"])
    doc = {"case": case}
    ns, model, params, keep, _ = case_inputs(doc)
    assert (params.market_mode == abi.MARKET_BOOTSTRAP) == variant.startswith("boot")
    if variant in ("boot_early", "boot_calm"):       # some pool is empty: synthetic Student-t blocks (bootstrap.py:214-222)
        assert min(pl[r].n_blocks for pl in (params.boot.pool_a, params.boot.pool_b) for r in range(2)) == 0
    G = params.sim_days + params.buffer_days
    n_rec = 12 if params.start_regime_prob1 > 0 else 3
    rec = _blank_tape(params, n_rec, 256)
    n = 2 * n_rec if params.antithetic_pairs else n_rec
    e = emu.run(params, rec, precision=abi.PREC_F64, rng_mode=abi.RNG_PHILOX_RECORD, seed=1234, n=n)
    if params.start_regime_prob1 > 0:
        assert 0 < rec.start_regime.sum() < n_rec and params.buffer_days == 0
    assert not (e.flags & abi.FLAG_SPELL_OVERFLOW).any()
    o = orc.run(params, rec)
    np.testing.assert_allclose(e.final_wealth, o.final_wealth, rtol=1e-11)
    np.testing.assert_array_equal(e.num_trades, o.num_trades)
    np.testing.assert_array_equal(e.start_offset, o.start_offset)
    np.testing.assert_array_equal(e.regime_days, o.regime_days)
    np.testing.assert_array_equal(e.jump_days, o.jump_days)
    # pure Philox mode (nothing recorded) is the same computation
    p = emu.run(params, HostTape(1, G, params.sim_days, 1), precision=abi.PREC_F64, rng_mode=abi.RNG_PHILOX,
                seed=1234, n=n)
    np.testing.assert_array_equal(p.final_wealth, e.final_wealth)


def test_philox_trials_are_pure_functions_of_seed_and_id():
    doc = {"case": dict(variant="default", years=0.5, n_tapes=1, seed=0, cfg={}, strategies=["S1", "S2", "S4", "S19"])}
    ns, model, params, keep, _ = case_inputs(doc)
    dummy = HostTape(1, params.sim_days + params.buffer_days, params.sim_days, 1)
    a = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=5, first=0, n=8)
    b = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=5, first=4, n=4)
    c = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=6, first=0, n=8)
    np.testing.assert_array_equal(a.final_wealth[:, 4:], b.final_wealth)
    np.testing.assert_array_equal(a.num_trades[:, 4:], b.num_trades)
    assert not np.array_equal(a.final_wealth, c.final_wealth)
    # antithetic twins share the start offset and regime path (mc_runner.py:145-153)
    np.testing.assert_array_equal(a.start_offset[0::2], a.start_offset[1::2])
    np.testing.assert_array_equal(a.regime_days[:, 0::2], a.regime_days[:, 1::2])


def _ks(sample, cdf, alpha=1e-4):
    d, p = stats.kstest(sample, cdf)
    assert p > alpha, (d, p)


@pytest.mark.parametrize("precision", [abi.PREC_F64, abi.PREC_F32])
def test_philox_draw_distributions(precision):
    """north star: on-device draws must match the reference's distributions (KS).  The recorded
    variates of the institutional generator: N(0,1), chi2(nu_regime), U[0,1), t(5), t(te_df)."""
    doc = {"case": dict(variant="vix_te", years=2, n_tapes=1, seed=0, cfg={}, strategies=["S1", "S2"])}
    ns, model, params, keep, _ = case_inputs(doc)
    params.asset_mask = 0x3F
    n_t = 24
    G, T = params.sim_days + params.buffer_days, params.sim_days
    rec = HostTape(n_t, G, T, 512)
    emu.run(params, rec, precision=precision, rng_mode=abi.RNG_PHILOX_RECORD, seed=99, n=2 * n_t)
    reg = rec.regime_full(params)
    _ks(rec.x.ravel(), stats.norm.cdf)
    # VIX / IRX noise is only consumed up to the end of the window (g < start_offset + T)
    used = np.arange(G)[:, None] < (rec.start_offset[None, :] + T)
    assert (rec.vix_n[~used] == 0).all()
    used[0] = False                                  # no VIX draw on generated day 0 (engine.py:103)
    _ks(rec.vix_n[used], stats.norm.cdf)
    _ks(rec.irx_n[used], stats.norm.cdf)
    for r in (0, 1):
        _ks(rec.chi2[reg == r], stats.chi2(params.joint[r].nu).cdf)
    _ks(rec.liq_n[1:].ravel(), stats.norm.cdf)
    _ks(rec.cred_n[1:].ravel(), stats.norm.cdf)
    _ks(rec.jump_u[1:].ravel(), stats.uniform.cdf)
    _ks(rec.jump_t[1:].ravel(), stats.t(5).cdf)
    assert abs(np.mean(rec.sign[1:] > 0) - 0.5) < 0.01
    for a in range(abi.N_ASSETS):       # df: 4.0, 6.5 (Marsaglia-Tsang), 7.0, 5, 10, 10
        _ks(rec.te_t[a, 1:].ravel(), stats.t(params.assets[a].te_df).cdf)
    # independence across streams / days: no correlation between the components of one day
    c = np.corrcoef(np.stack([rec.x[:, 0].ravel(), rec.x[:, 1].ravel(), rec.x[:, 2].ravel(),
                              rec.chi2.ravel(), np.where(used, rec.vix_n, rec.x[:, 0] * 0 + np.random.default_rng(1).standard_normal(used.shape)).ravel()]))
    assert np.abs(c - np.eye(5)).max() < 0.01
    # dwell times: geometric with p = 1 - clip(p_stay, .80, .995)
    sp0 = rec.spells[0::2][:8].ravel().astype(float)
    _ks(sp0 - np.random.default_rng(0).random(sp0.size), stats.expon(scale=-1 / np.log(0.99)).cdf, alpha=1e-5)
    so = rec.start_offset
    assert so.min() >= 50 and so.max() < 1260


def _c2_params(sids, years=2, **assets_override):
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                    USE_ANTITHETIC_VARIATES=False)
    for a, kv in assets_override.items():
        ns.ASSETS[a].update(kv)
    ns.LETF_CONFIGS = ns.ASSETS
    return build_params(sids, years, synthetic.normal_regime_model("c2"), ns)


def test_lean_fast_variant_matches_generic_build():
    """The loop-unswitched straight-line day step of the lean buy&hold build (what bench.py's c2 workload
    runs) against the generic build on the same Philox counters: same draws, fp32 reassociation only."""
    sids = ["S1", "B_UPRO", "S3"]
    params, keep = _c2_params(sids)
    dummy = HostTape(1, params.sim_days, params.sim_days, 1)
    lean, fast = emu.run_lean(params, 64, seed=77)
    assert fast
    gen = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=77, n=64)
    np.testing.assert_allclose(lean.final_wealth, gen.final_wealth, rtol=2e-5)
    np.testing.assert_allclose(lean.max_dd, gen.max_dd, atol=2e-6)
    np.testing.assert_array_equal(lean.flags, gen.flags)
    # the lean build's general loop (what runs when the shape is not admitted) on the same counters
    slow, _ = emu.run_lean(params, 64, seed=77, force_generic=True)
    np.testing.assert_allclose(lean.final_wealth, slow.final_wealth, rtol=2e-5)
    # the launch form (two interleaved trials per thread, odd tail) computes the same thing
    two, fast2 = emu.run_lean(params, 63, seed=77, two_per_thread=True)
    assert fast2
    np.testing.assert_allclose(two.final_wealth, lean.final_wealth[:, :63], rtol=1e-6)
    np.testing.assert_allclose(two.max_dd, lean.max_dd[:, :63], atol=1e-6)
    np.testing.assert_array_equal(two.flags, lean.flags[:63])
    np.testing.assert_array_equal(two.regime_days, lean.regime_days[:, :63])
    # and against fp64 on the same counters (fp64 Box-Muller / log differ in the last bits of each variate)
    g64 = emu.run(params, dummy, precision=abi.PREC_F64, rng_mode=abi.RNG_PHILOX, seed=77, n=64)
    np.testing.assert_allclose(lean.final_wealth, g64.final_wealth, rtol=5e-3)
    # non-default tracking-error dof -> the lean build must fall back to its general loop, same results
    p2, k2 = _c2_params(sids, UPRO={"tracking_error_df": 7})
    lean2, fast2 = emu.run_lean(p2, 16, seed=78)
    assert not fast2
    gen2 = emu.run(p2, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=78, n=16)
    np.testing.assert_allclose(lean2.final_wealth, gen2.final_wealth, rtol=2e-5)


@pytest.mark.parametrize("sids,code", [(["S2", "B_QQQ", "B_TLT"], 0x203), (["S1", "S2", "S3"], 0x304), (["S1"], 0x100), (["S2", "B_TLT"], 0x101)])
def test_joint_pair_variant_matches_generic_build(sids, code):
    """One antithetic pair per thread (letf_jfast.cuh, what bench.py's c3 workload runs): the twins share every Philox block
    and variate, as the reference's identically seeded twins do (mc_runner.py:145-153), against the lean build's general
    loop and the tape-capable generic build on the same counters: same draws, fp32 reassociation only."""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, 2, synthetic.fallback_regime_model("default"), ns)
    pair, used = emu.run_jpair(params, 48, seed=11)
    assert used == code
    slow, _ = emu.run_jpair(params, 48, seed=11, force_generic=True)
    dummy = HostTape(1, params.sim_days + params.buffer_days, params.sim_days, 1)
    gen = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=11, n=48)
    for ref in (slow, gen):
        np.testing.assert_allclose(pair.final_wealth, ref.final_wealth, rtol=2e-5)
        np.testing.assert_allclose(pair.max_dd, ref.max_dd, atol=5e-6)
        np.testing.assert_array_equal(pair.flags, ref.flags)
        np.testing.assert_array_equal(pair.start_offset, ref.start_offset)
        np.testing.assert_array_equal(pair.regime_days, ref.regime_days)
        np.testing.assert_array_equal(pair.jump_days, ref.jump_days)
    assert pair.jump_days.sum() > 0 and (pair.regime_days[1] > 0).any()
    # twins share the start offset and the regime path, and differ
    np.testing.assert_array_equal(pair.start_offset[0::2], pair.start_offset[1::2])
    assert not np.allclose(pair.final_wealth[:, 0::2], pair.final_wealth[:, 1::2])
    # a trial range that starts on an odd twin and ends on an even one
    part, _ = emu.run_jpair(params, 9, seed=11, first=3)
    np.testing.assert_array_equal(part.final_wealth, pair.final_wealth[:, 3:12])
    np.testing.assert_array_equal(part.jump_days, pair.jump_days[3:12])


def test_joint_pair_variant_not_admitted_shapes_fall_back():
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_ANTITHETIC_VARIATES=False)
    params, keep = build_params(["S2", "B_QQQ", "B_TLT"], 1, synthetic.fallback_regime_model("default"), ns)
    out, used = emu.run_jpair(params, 4, seed=1)
    assert used == 0 and np.isfinite(out.final_wealth).all()
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    # a calibrated tracking_residual_model (TQQQ: dof 4 + tanh cap) runs the pair body with the run-time sampler (code 0x80)
    params, keep = build_params(["S1", "S2"], 1, synthetic.fallback_regime_model("vix_te"), ns)
    pair, used = emu.run_jpair(params, 16, seed=1)
    slow, _ = emu.run_jpair(params, 16, seed=1, force_generic=True)
    assert used == 0x280
    np.testing.assert_allclose(pair.final_wealth, slow.final_wealth, rtol=2e-5)
    np.testing.assert_array_equal(pair.jump_days, slow.jump_days)
    # explicit VIX dynamics, empirical dwell-time samples and non-integer joint dof (calibrated models, calibration.py:399:
    # Marsaglia-Tsang chi-square with out-of-line retries) are run-time parameters of the pair kernel
    for variant in ("vix_te", "durations", "jumpy", "calibrated_nu"):
        model = synthetic.fallback_regime_model("default" if variant == "calibrated_nu" else variant)
        if variant == "calibrated_nu":
            model["joint_return_model"]["regimes"][0]["nu"] = 5.3
            model["joint_return_model"]["regimes"][1]["nu"] = 4.2
        params, keep = build_params(["S2", "B_QQQ", "B_TLT"], 1, model, ns)
        pair, used = emu.run_jpair(params, 16, seed=2)
        slow, _ = emu.run_jpair(params, 16, seed=2, force_generic=True)
        assert used == 0x203
        np.testing.assert_allclose(pair.final_wealth, slow.final_wealth, rtol=2e-5)
        np.testing.assert_array_equal(pair.regime_days, slow.regime_days)
        np.testing.assert_array_equal(pair.jump_days, slow.jump_days)
        np.testing.assert_array_equal(pair.flags, slow.flags)


def test_philox_distribution_matches_live_reference():
    """north star: 'On-device Philox draws must match the reference's distributions by KS test and
    summary-quantile agreement' - the reference side is the LIVE reference with its own PCG64 streams
    (2 400 trials, institutional_v1 defaults, 2-year horizon, 10 strategies)."""
    from tests.common import assert_same_distribution, reference_distribution
    ref = reference_distribution()
    sids = [str(s) for s in ref["strategies"]]
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, float(ref["years"]), synthetic.fallback_regime_model("default"), ns)
    dummy = HostTape(1, params.sim_days + params.buffer_days, params.sim_days, 1)
    out = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=11, n=8000)
    assert_same_distribution(ref, out, sids)


def test_philox_distribution_matches_live_reference_at_10_years():
    """the longer fixture (tests/golden/ref_distribution_10y.npz: 2 000 trials x 10 years x 10 strategies of the unmodified
    reference with its own NumPy streams): two-sample KS p > 0.01 on terminal wealth and max drawdown per strategy, the
    reference's P10 ... P90 within +- 3 rank points of ours, trade counts within 3 %"""
    from tests.common import assert_same_distribution, reference_distribution
    ref = reference_distribution("ref_distribution_10y.npz")
    sids = [str(s) for s in ref["strategies"]]
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    params, keep = build_params(sids, float(ref["years"]), synthetic.fallback_regime_model("default"), ns)
    dummy = HostTape(1, params.sim_days + params.buffer_days, params.sim_days, 1)
    out = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=11, n=6000)
    assert_same_distribution(ref, out, sids, p_min=0.01, rank_tol=3.0)


def test_bootstrap_philox_distribution_matches_live_reference():
    """same for the legacy_hybrid block bootstrap: the reference side ran its own BlockBootstrapReturns with NumPy's
    geometric / integers / choice(p=...) / standard_t streams (tests/golden/ref_distribution_boot_1y.npz)."""
    from tests.common import assert_same_distribution, bootstrap_distribution_case
    ref, sids, params, keep = bootstrap_distribution_case()
    assert params.market_mode == abi.MARKET_BOOTSTRAP
    out = emu.run(params, _blank_tape(params, 1, 1), precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=23, n=8000)
    assert_same_distribution(ref, out, sids)


def test_historical_anchor_start_matches_live_reference():
    """'historical_anchor' start: Bernoulli start regime here, a uniformly drawn history row in the reference
    (tests/golden/ref_distribution_anchor_1y.npz: 4 000 trials, institutional_v1, 1 year, no buffer)."""
    from scipy import stats as st
    from tests.common import anchor_distribution_case, assert_same_distribution
    ref, sids, params, keep = anchor_distribution_case()
    assert params.buffer_days == 0 and 0.05 < params.start_regime_prob1 < 0.95
    out = emu.run(params, HostTape(1, params.sim_days, params.sim_days, 1), precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX,
                  seed=29, n=8000)
    assert_same_distribution(ref, out, sids)
    # the start regime drives the regime-day counts of a 1-year window directly: compare their whole distribution
    assert st.ks_2samp(ref["high_vol_days"], out.regime_days[1]).pvalue > 1e-3


@pytest.mark.parametrize("slot_list,sids", [(1, ["S1", "S2", "S4", "S5", "S6", "S19"]), (2, ALL19)])
def test_typed_slot_builds_match_dynamic_dispatch(slot_list, sids):
    """one compile-time rule per slot + compile-time asset mask (the c4 and default-sweep builds) must
    compute exactly what the dynamically dispatched generic build computes on the same Philox counters"""
    ns = synthetic.config_namespace()
    params, keep = build_params(sids, 2, synthetic.fallback_regime_model("default"), ns)
    dummy = HostTape(1, params.sim_days + params.buffer_days, params.sim_days, 1)
    typed = emu.run_typed(params, slot_list, 48, seed=5)
    gen = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=5, n=48)
    np.testing.assert_array_equal(typed.final_wealth, gen.final_wealth)
    np.testing.assert_array_equal(typed.max_dd, gen.max_dd)
    np.testing.assert_array_equal(typed.num_trades, gen.num_trades)
    np.testing.assert_array_equal(typed.flags, gen.flags)
    np.testing.assert_array_equal(typed.regime_days, gen.regime_days)


@pytest.mark.parametrize("slot_list,sids", [(1, ["S1", "S2", "S4", "S5", "S6", "S19"]), (2, ALL19)])
def test_split_pipeline_matches_fused_typed_build(slot_list, sids):
    """The split strategy pipeline (what the c4 / default-sweep shapes run): the antithetic-pair producer writes per-day
    records (layer returns, VIX + regime, SPY price), the typed rule kernel body consumes them.  Same Philox counters as
    the fused typed build: market state identical, wealth to fp32 reassociation, trade decisions almost everywhere equal."""
    ns = synthetic.config_namespace()
    params, keep = build_params(sids, 2, synthetic.fallback_regime_model("default"), ns)
    split = emu.run_split(params, slot_list, 48, seed=5)
    fused = emu.run_typed(params, slot_list, 48, seed=5)
    # calibrated tracking_residual_model (non-integer dof, tanh cap): the producer with the run-time sampler
    pv, keepv = build_params(sids, 1, synthetic.fallback_regime_model("vix_te"), ns)
    dummy = HostTape(1, pv.sim_days + pv.buffer_days, pv.sim_days, 1)
    sv = emu.run_split(pv, slot_list, 16, seed=6)
    fv = emu.run(pv, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=6, n=16)      # generic build: run-time dispatch
    np.testing.assert_allclose(sv.final_wealth, fv.final_wealth, rtol=1e-3)
    assert np.mean(sv.num_trades != fv.num_trades) < 0.02
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=1e-3)
    assert np.median(np.abs(split.final_wealth / fused.final_wealth - 1)) < 2e-6
    np.testing.assert_allclose(split.max_dd, fused.max_dd, atol=5e-5)
    assert np.mean(split.num_trades != fused.num_trades) < 0.01
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(split, k), getattr(fused, k))
    # an odd number of trials (the last pair's second twin is computed and not stored)
    part = emu.run_split(params, slot_list, 7, seed=5)
    np.testing.assert_array_equal(part.final_wealth, split.final_wealth[:, :7])
    np.testing.assert_array_equal(part.num_trades, split.num_trades[:, :7])
    # a launch that starts on an odd trial: the first pair's even twin is computed, its records written, nothing of it stored
    off = emu.run_split(params, slot_list, 6, seed=5, first=3)
    np.testing.assert_array_equal(off.final_wealth, split.final_wealth[:, 3:9])
    np.testing.assert_array_equal(off.num_trades, split.num_trades[:, 3:9])
    np.testing.assert_array_equal(off.flags, split.flags[3:9])


@pytest.mark.parametrize("sids,extra,model,years", [
    (["S1", "S4", "S7", "S10", "S15"], False, "default", 2),              # a subset of the rules, no typed list for it
    (["S2", "S3", "B_QQQ", "B_UPRO", "B_TLT", "S9", "S13"], True, "default", 2),     # all six series incl. QQQ
    (["S2", "S3", "B_QQQ", "B_UPRO", "B_TLT", "S9", "S13", "S4", "S5", "S7"], True, "default", 1),      # more than eight: 32-slot build
    (["S1", "S3", "S2", "B_UPRO", "B_TMF"], True, "default", 1),          # buy&hold only, more series than the pair kernels carry
    (["S4", "S6", "S12", "S18"], False, "vix_te", 1),                      # calibrated tracking-error model: run-time sampler
])
def test_generic_split_pipeline_matches_fused_generic_build(sids, extra, model, years):
    """Strategy lists without a typed build run the generic split pipeline: the uncompacted pair producer (slot = asset,
    series outside the mask skipped and written as zeros) -> six-field records -> the dynamic-slot rule kernel body.
    Same Philox counters as the fused generic build: market state identical, wealth to fp32 reassociation."""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS) if extra else synthetic.config_namespace()
    params, keep = build_params(sids, years, synthetic.fallback_regime_model(model), ns)
    dummy = HostTape(1, params.sim_days + params.buffer_days, params.sim_days, 1)
    split = emu.run_split(params, 0, 24, seed=8)
    fused = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=8, n=24)
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=1e-3)
    assert np.median(np.abs(split.final_wealth / fused.final_wealth - 1)) < 5e-6
    np.testing.assert_allclose(split.max_dd, fused.max_dd, atol=5e-5)
    assert np.mean(split.num_trades != fused.num_trades) < 0.02
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(split, k), getattr(fused, k))
    part = emu.run_split(params, 0, 5, seed=8)
    np.testing.assert_array_equal(part.final_wealth, split.final_wealth[:, :5])


@pytest.mark.parametrize("sids,extra,model,lst", [
    (["S1", "S4", "S7", "S10", "S15"], False, "default", 3),
    (["S19", "S2", "S8", "S6"], False, "default", 3),                               # out of list order; S8 reads SPY
    (["S3", "S1", "S2", "B_TMF", "B_TLT", "S12"], True, "default", 3),             # more buy&hold strategies than the list's three slots
    (["S5", "S9", "S17"], False, "vix_te", 3),
    (["S4", "S2"], False, "default", 4),                                            # subsets of the config-4 list: its six-slot kernel
    (["S19", "S5", "B_TMF", "S1", "S2", "B_TLT"], True, "default", 4),
    (["S6", "S4"], False, "vix_te", 4),
])
def test_routed_split_pipeline_matches_fused_generic_build(sids, extra, model, lst):
    """Any subset of the default sweep runs the 19-rule kernel with the other slots switched off and the outputs routed to
    the request's rows (slot list 3); the producer simulates the list's four series.  Against the fused generic build on
    the same Philox counters."""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS) if extra else synthetic.config_namespace()
    params, keep = build_params(sids, 2, synthetic.fallback_regime_model(model), ns)
    dummy = HostTape(1, params.sim_days + params.buffer_days, params.sim_days, 1)
    split = emu.run_split(params, lst, 24, seed=8)
    fused = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=8, n=24)
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=1e-3)
    assert np.median(np.abs(split.final_wealth / fused.final_wealth - 1)) < 5e-6
    np.testing.assert_allclose(split.max_dd, fused.max_dd, atol=5e-5)
    assert np.mean(split.num_trades != fused.num_trades) < 0.02
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(split, k), getattr(fused, k))
    # a request the list cannot hold: the same rule type twice / a rule on a series outside the list
    two_sma, _ = build_params(["S4", "S4"], 1, synthetic.fallback_regime_model("default"), ns)
    with pytest.raises(RuntimeError):
        emu.run_split(two_sma, 3, 2, seed=1)


@pytest.mark.parametrize("days,random_start", [(2, True), (5, False), (61, True), (257, False), (300, True)])
def test_pair_and_split_paths_at_edge_horizons(days, random_start):
    """horizons around the staging depth (3 days ahead), the window lengths and the steady-state switch of the day body
    (256), with and without the random-start buffer: pair kernel and split pipeline against the general loops"""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=random_start)
    model = synthetic.fallback_regime_model("default")
    years = days / 252.0 + 1e-9
    params, keep = build_params(["S2", "B_QQQ", "B_TLT"], years, model, ns)
    assert params.sim_days == days
    pair, used = emu.run_jpair(params, 10, seed=3)
    slow, _ = emu.run_jpair(params, 10, seed=3, force_generic=True)
    assert used == 0x203
    np.testing.assert_allclose(pair.final_wealth, slow.final_wealth, rtol=2e-5)
    np.testing.assert_allclose(pair.max_dd, slow.max_dd, atol=5e-6)
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(pair, k), getattr(slow, k))
    p19, keep19 = build_params(ALL19, years, model, synthetic.config_namespace(USE_RANDOM_START=random_start))
    split = emu.run_split(p19, 2, 10, seed=4)
    fused = emu.run_typed(p19, 2, 10, seed=4)
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=1e-3)
    np.testing.assert_array_equal(split.regime_days, fused.regime_days)
    np.testing.assert_array_equal(split.flags, fused.flags)
    assert np.mean(split.num_trades != fused.num_trades) < 0.02


def test_lean_fast_variant_config1_shape():
    """BASELINE config 1 (validation.py:33-57): one frictionless 3x series - the straight-line loop draws no
    tracking-error block at all; against the lean build's general loop and the generic build on the same counters"""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                    USE_ANTITHETIC_VARIATES=False, LETF_CONFIGS=synthetic.pure_leverage_assets())
    ns.ASSETS = ns.LETF_CONFIGS
    params, keep = build_params(["B_UPRO"], 1, synthetic.normal_regime_model("c1"), ns)
    fast, code = emu.run_lean(params, 33, seed=3)
    assert code == 1 | (1 << 8) | (1 << 12)
    two, _ = emu.run_lean(params, 33, seed=3, two_per_thread=True)
    slow, _ = emu.run_lean(params, 33, seed=3, force_generic=True)
    dummy = HostTape(1, params.sim_days, params.sim_days, 1)
    gen = emu.run(params, dummy, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=3, n=33)
    for other in (two, slow, gen):
        np.testing.assert_allclose(fast.final_wealth, other.final_wealth, rtol=2e-6)
        np.testing.assert_allclose(fast.max_dd, other.max_dd, atol=1e-6)
    # frictionless: terminal wealth is exactly the compounded 3x underlying, so log growth has mean -L^2 sigma^2 / 2 per year
    big, _ = emu.run_lean(params, 4000, seed=9, two_per_thread=True)
    ann = np.median(big.final_wealth[0] / params.initial_capital - 1.0)
    assert abs(ann - (-0.10125)) < 0.015


def test_lean_fast_variant_config5_shape():
    """BASELINE config 5 cell: L = 3 / 2 / 1 on one SPY-like underlying (UPRO, SSO, SPY rows of cfg.ASSETS;
    SPY's tracking error is t(10)) -> the "last slot t(10)" straight-line loop against the general loop"""
    sids = ["B_UPRO", "S3", "S2"]
    params, keep = _c2_params(sids, years=2)
    fast, code = emu.run_lean(params, 33, seed=3)
    assert code == (1 + 4) | (3 << 8)          # slots UPRO, SSO, SPY: third one is t(10)
    two, _ = emu.run_lean(params, 33, seed=3, two_per_thread=True)
    slow, _ = emu.run_lean(params, 33, seed=3, force_generic=True)
    np.testing.assert_allclose(fast.final_wealth, slow.final_wealth, rtol=2e-5)
    np.testing.assert_allclose(two.final_wealth, fast.final_wealth, rtol=1e-6)
    np.testing.assert_allclose(fast.max_dd, slow.max_dd, atol=2e-6)


def test_bootstrap_block_draws_follow_the_reference_rules():
    """Philox block draws of the legacy_hybrid bootstrap (bootstrap.py:302-308, 212-262): recorded block lengths are
    geometric(1 / BLOCK_MEAN), the six selection variates uniform, both noise matrices t(STUDENT_T_DF); and the
    stitched market is a pure function of (seed, trial): a second run reproduces it."""
    doc = {"case": dict(variant="boot", years=2, n_tapes=1, seed=0, cfg={"SIM_ENGINE_MODE": "legacy_hybrid"},
                        strategies=["S1", "S2", "B_QQQ", "B_TLT"])}
    ns, model, params, keep, _ = case_inputs(doc)
    n_t = 64
    rec = _blank_tape(params, n_t, 512)
    a = emu.run(params, rec, precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX_RECORD, seed=11, n=2 * n_t)
    b = emu.run(params, _blank_tape(params, 1, 1), precision=abi.PREC_F32, rng_mode=abi.RNG_PHILOX, seed=11, n=2 * n_t)
    np.testing.assert_array_equal(a.final_wealth, b.final_wealth)
    G = rec.gen_days
    lens = rec.boot_len.astype(np.int64)
    used = np.zeros_like(lens, dtype=bool)                 # blocks actually consumed: until the column covers G days
    cover = np.zeros(n_t, np.int64)
    for k in range(rec.max_blocks):
        used[k] = cover < G
        cover += np.clip(lens[k], params.boot.block_min, params.boot.block_max)
    p = 1.0 / params.boot.block_mean
    _ks(lens[used] - 1 + np.random.default_rng(0).random(used.sum()), lambda x: 1 - (1 - p) ** np.maximum(x, 0))  # jittered geometric
    for s in range(6):
        _ks(rec.boot_u[:, s, :][used], stats.uniform.cdf)
    cons = np.arange(G)[:, None] < (rec.start_offset[None, :] + params.sim_days)
    for c in range(3):
        _ks(rec.boot_t[:, c, :][cons], stats.t(params.param_t_df).cdf)
        _ks(rec.x[:, c, :][cons], stats.t(params.param_t_df).cdf)


def _reference_frame():
    import os
    from oracle.make_strategy_golden import strategy_frame
    from tests.common import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "run_strategy_fixed_700d.npz"))
    df, regime = strategy_frame(int(g["T"]), int(g["seed"]))
    return g, df, regime


def test_series_mode_equity_curves_match_live_reference():
    """materialised-path mode of the kernel source (run_strategy_fixed seam) on the frame the LIVE reference's
    run_strategy_fixed was run on (tests/golden/run_strategy_fixed_700d.npz): daily equity curve to 1e-11 at every day,
    identical trade counts, journal entry counts and journal days, for 13 rules."""
    from letf_monte_carlo_b200.trade import TAXABLE_IDS
    g, df, regime = _reference_frame()
    T = len(df)
    frame = {a: df[f"{a}_Ret"].to_numpy().reshape(T, 1) for a in abi.ASSET_NAMES}
    ns = synthetic.config_namespace()
    for sid in [str(x) for x in g["sids"]]:
        params, keep = build_params([sid], T / 252.0, synthetic.fallback_regime_model(), ns)
        out = emu.run_series(params, frame, tlt_ret=df["TLT_Ret"].to_numpy().reshape(T, 1), vix=df["VIX"].to_numpy().reshape(T, 1),
                             regime=regime.reshape(T, 1), log=(0, 0, 4096), curve=True)
        np.testing.assert_allclose(out.curve, g[f"{sid}_curve"], rtol=1e-11, err_msg=sid)
        assert out.num_trades[0, 0] == int(g[f"{sid}_trades"]), sid
        assert out.final_wealth[0, 0] == out.curve[-1]
        if sid in TAXABLE_IDS:
            n = int(out.log["count"][0])
            assert n == int(g[f"{sid}_journal_n"]), sid
            np.testing.assert_array_equal(out.log["day"][:n], g[f"{sid}_journal_days"])


class _PresetRng:
    """stands in for the numpy Generator handed to simulate_joint_returns_t: returns pre-drawn variates"""

    def __init__(self, x, chi2):
        self.x, self.chi2 = x, chi2

    def standard_normal(self, size=None):
        assert tuple(size) == self.x.shape
        return self.x

    def chisquare(self, df):
        assert len(df) == len(self.chi2)
        return self.chi2


@pytest.mark.parametrize("mm", [True, False])
@pytest.mark.parametrize("anti", [False, True])
def test_joint_returns_seam_matches_live_reference(mm, anti):
    """simulate_joint_returns_t(n_days, regime_path, joint_model, rng, antithetic) (calibration.py:480-590) along a caller's
    regime path: the kernel source fed with the drop-in's parameter block and tape, against the LIVE reference on the same
    variates (tests/golden/joint_returns_400d.npz): every daily return of the three series to 1e-12, clips included"""
    import os
    from letf_monte_carlo_b200.calibration import joint_case
    from oracle.make_joint_returns_golden import inputs
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "joint_returns_400d.npz"))
    regime, x, chi2, joint = inputs()
    np.testing.assert_array_equal(x, g["x"])
    ns = synthetic.config_namespace(USE_MOMENT_MATCHING=mm)
    params, keep, tape, trial = joint_case(len(regime), regime, joint, _PresetRng(x, chi2), anti, ns)
    ret = emu.market_returns(params, tape, trial)
    want = g[f"ret_mm{int(mm)}_anti{int(anti)}"]
    np.testing.assert_allclose(ret, want, rtol=1e-12, atol=1e-15)
    assert np.abs(want).max() > 0.9          # the clip at -0.95 is exercised


def test_series_mode_with_inferred_regime_path_matches_live_reference():
    """run_strategy_fixed(df, sid, None, ...): the regime path is inferred from the frame (strategy.py:64-78); kernel source in
    series mode fed with `regime_path_for_frame` against the live reference's curves (tests/golden/infer_regime_900d.npz)"""
    import os
    from letf_monte_carlo_b200.utils import regime_path_for_frame
    from oracle.make_regime_inference_golden import frame as mk
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "infer_regime_900d.npz"))
    df = mk()
    df["VIX"] = df["VIX"].ffill()
    T = len(df)
    frame = {a: df[f"{a}_Ret"].to_numpy().reshape(T, 1) for a in abi.ASSET_NAMES}
    ns = synthetic.config_namespace()
    for sid in ("S4", "S8", "S16"):
        params, keep = build_params([sid], T / 252.0, synthetic.fallback_regime_model(), ns)
        for given, key in ((None, f"{sid}_curve"), (np.zeros(17, dtype=int), f"{sid}_curve_mismatch")):
            reg = regime_path_for_frame(df, given).astype(np.int32)
            out = emu.run_series(params, frame, tlt_ret=df["TLT_Ret"].to_numpy().reshape(T, 1), vix=df["VIX"].to_numpy().reshape(T, 1),
                                 regime=reg.reshape(T, 1), cash_ret=df["Cash_Ret"].to_numpy().reshape(T, 1), log=(0, 0, 4096),
                                 curve=True)
            np.testing.assert_allclose(out.curve, g[key], rtol=1e-11, err_msg=sid)
        assert out.num_trades[0, 0] == int(g[f"{sid}_trades"]), sid


@pytest.mark.parametrize("T", [1, 2, 3, 5, 61, 203])
def test_series_mode_short_frames_and_typed_build(T):
    """edge lengths of the series prefetch queue (T below / around its depth and the window lengths), generic fp64 and
    fp32 builds and the typed 19-rule fp32 build against the oracle's run_strategy on the same frame"""
    import ctypes as C
    rng = np.random.default_rng(100 + T)
    n = 5
    ns = synthetic.config_namespace()
    params, keep = build_params(ALL19, max(T, 1) / 252.0, synthetic.fallback_regime_model(), ns)
    rets = {a: np.ascontiguousarray(rng.standard_t(5, (T, n)) * s) for a, s in
            zip(abi.ASSET_NAMES, (0.03, 0.03, 0.02, 0.025, 0.01, 0.013))}
    vix = np.ascontiguousarray(15 + 10 * rng.random((T, n)))
    regime = np.ascontiguousarray((rng.random((T, n)) < 0.3).astype(np.int32))
    tlt = np.ascontiguousarray(rng.normal(0, 0.008, (T, n)))
    o = orc.HostOutputs(params.n_strategies, n)
    s = abi.Series()
    s.n_paths, s.sim_days = n, T
    for a, name in enumerate(abi.ASSET_NAMES):
        s.asset_ret[a] = rets[name].ctypes.data
    s.tlt_ret, s.vix, s.regime = tlt.ctypes.data, vix.ctypes.data, regime.ctypes.data
    assert orc.lib().letf_oracle_run_strategies(C.byref(params), C.byref(s), C.byref(o.as_struct())) == 0
    e64 = emu.run_series(params, rets, tlt_ret=tlt, vix=vix, regime=regime, precision=abi.PREC_F64)
    np.testing.assert_allclose(e64.final_wealth, o.final_wealth, rtol=1e-12)
    np.testing.assert_array_equal(e64.num_trades, o.num_trades)
    e32 = emu.run_series(params, rets, tlt_ret=tlt, vix=vix, regime=regime, precision=abi.PREC_F32)
    typed = emu.run_series(params, rets, tlt_ret=tlt, vix=vix, regime=regime, precision=abi.PREC_F32, typed=True)
    np.testing.assert_allclose(e32.final_wealth, o.final_wealth, rtol=5e-4)
    np.testing.assert_allclose(typed.final_wealth, e32.final_wealth, rtol=1e-4)
    assert np.mean(typed.num_trades != e32.num_trades) < 0.02


@pytest.mark.parametrize("T", [1, 2, 5, 61, 203, 700])
def test_series_mode_through_the_record_pass(T):
    """materialised-path mode through the split pipeline (letf_series_prep.cuh: record pass over the caller's series, then the
    rule kernel body in record mode) - typed, routed and dynamic-slot rule kernels against the fused series builds on the same
    frame: buy&hold rows, flags and regime counts bit for bit (same arithmetic), rule rows to fp32 reassociation"""
    rng = np.random.default_rng(300 + T)
    n = 6
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
    rets = {a: np.ascontiguousarray(rng.standard_t(5, (T, n)) * s) for a, s in
            zip(abi.ASSET_NAMES, (0.03, 0.03, 0.02, 0.025, 0.01, 0.013))}
    vix = np.ascontiguousarray(15 + 10 * rng.random((T, n)))
    vix[T // 2, 0] = 130.0                                  # out of the validated band on one path -> LETF_FLAG_REPAIRED
    regime = np.ascontiguousarray((rng.random((T, n)) < 0.3).astype(np.int32))
    tlt = np.ascontiguousarray(rng.normal(0, 0.008, (T, n)))
    years = max(T, 1) / 252.0
    cases = [(ALL19, 2), (["S19", "S2", "S8", "S6", "B_TLT"], 3), (["S4", "S9", "B_QQQ", "B_UPRO", "S2"], 0)]
    for sids, lst in cases:
        params, keep = build_params(sids, years, synthetic.fallback_regime_model(), ns)
        fused = emu.run_series(params, rets, tlt_ret=tlt, vix=vix, regime=regime, precision=abi.PREC_F32, typed=(lst == 2))
        split = emu.run_series(params, rets, tlt_ret=tlt, vix=vix, regime=regime, split_list=lst)
        hold = [k for k, sid in enumerate(sids) if ns.STRATEGIES[sid]["type"].startswith("benchmark")]
        np.testing.assert_array_equal(split.final_wealth[hold], fused.final_wealth[hold])
        np.testing.assert_array_equal(split.max_dd[hold], fused.max_dd[hold])
        for k in ("flags", "start_offset", "regime_days", "jump_days"):
            np.testing.assert_array_equal(getattr(split, k), getattr(fused, k))
        assert split.flags[0] & abi.FLAG_REPAIRED
        np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=2e-4)
        np.testing.assert_allclose(split.max_dd, fused.max_dd, atol=5e-5)
        assert np.mean(split.num_trades != fused.num_trades) < 0.03
        # a frame with its own Cash_Ret column (what the reference's frames carry, strategy.py:138): copied beside the records
        cash = np.ascontiguousarray(2e-4 + 1e-4 * rng.random((T, n)))
        fc = emu.run_series(params, rets, tlt_ret=tlt, vix=vix, regime=regime, cash_ret=cash, precision=abi.PREC_F32, typed=(lst == 2))
        sc = emu.run_series(params, rets, tlt_ret=tlt, vix=vix, regime=regime, cash_ret=cash, split_list=lst)
        np.testing.assert_allclose(sc.final_wealth, fc.final_wealth, rtol=2e-4)
        if T > 5 and lst != 0:
            assert np.abs(sc.final_wealth / split.final_wealth - 1).max() > 1e-7        # the column is really read


def test_pair_kernel_and_split_pipeline_without_moment_matching():
    """USE_MOMENT_MATCHING = False (calibration.py:579): the pair kernel and the split pipeline keep their two passes and shift
    by zero - same Philox counters as the general one-pass loops"""
    ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_MOMENT_MATCHING=False)
    model = synthetic.fallback_regime_model("default")
    params, keep = build_params(["S2", "B_QQQ", "B_TLT"], 1, model, ns)
    assert not params.moment_matching
    pair, used = emu.run_jpair(params, 12, seed=3)
    slow, _ = emu.run_jpair(params, 12, seed=3, force_generic=True)
    assert used == 0x203
    np.testing.assert_allclose(pair.final_wealth, slow.final_wealth, rtol=2e-5)
    np.testing.assert_allclose(pair.max_dd, slow.max_dd, atol=5e-6)
    for k in ("flags", "start_offset", "regime_days", "jump_days"):
        np.testing.assert_array_equal(getattr(pair, k), getattr(slow, k))
    p19, keep19 = build_params(ALL19, 1, model, synthetic.config_namespace(USE_MOMENT_MATCHING=False))
    split = emu.run_split(p19, 2, 12, seed=4)
    fused = emu.run_typed(p19, 2, 12, seed=4)
    np.testing.assert_allclose(split.final_wealth, fused.final_wealth, rtol=1e-3)
    np.testing.assert_array_equal(split.regime_days, fused.regime_days)
    np.testing.assert_array_equal(split.flags, fused.flags)
    # and it is NOT the moment-matched result
    pmm, _ = build_params(["S2", "B_QQQ", "B_TLT"], 1, model, synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS))
    mm, _ = emu.run_jpair(pmm, 12, seed=3)
    assert np.abs(mm.final_wealth / pair.final_wealth - 1).max() > 1e-4
