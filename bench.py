#!/usr/bin/env python
"""bench.py - LETF path-days/sec on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c1x|c2|c3|c4|c5|sweep19|sub5|sma2|boot|series|series19]

A "step" is one pass of the hot path over one batch: the path kernel(s) over this rank's trials (Philox draws generated on
the device), the summary reduction kernel, and - for N > 1 - the exchange of the packed summary (one NCCL all-gather + one
combine kernel).  N = 1 workload: BASELINE.json configs[2], the largest single-GPU configuration ("c3": regime-switching
multivariate-t GARCH/DCC QQQ/SPY/TLT market with VIX / IRX, stress + jumps, antithetic pairs, moment matching, random
start; 1M trials x 7,560 days).  N > 1: the same number of trials PER GPU (weak scaling), trial ids globally unique.

The line carries `roofline` (ALU-issue roofline of SURVEY.md 8d, measured DRAM traffic vs algorithmic bytes, measured pipe
peaks), `e2e` (the same metric through the public API: parameter dicts in, summary table out on the host),
`cpu_baseline` (the C port of the reference algorithm on the box's host cores) and `cpu_baseline_numpy` (the UNMODIFIED
reference package, copied to the git-ignored baseline/_ref, through its own parallel_monte_carlo_fixed).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/letf_oracle.c, every variate drawn by a
conventional CPU generator) with all host threads, on a bounded sample of the same workload - the conservative
denominator: the reference's own NumPy path is three orders of magnitude slower.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "LETF path-days/sec"
UNIT = "path-days/s"

# algorithmic lane-instructions per path-day (Philox4x32-10, inverse-CDF / Box-Muller
# normals, gamma-based chi2 / Student-t, compounding + drawdown per tracked equity); restated in DESIGN.md
WORKLOADS = {
    "c1": dict(desc="BASELINE configs[0]: TQQQ-like 3x daily-reset single-asset GBM (validation.py:33-57: zero drift, sigma 15 %/yr, "
                    "no frictions), buy&hold, 10k trials x 252 days (literal size: launch-bound)",
               trials=10_000, years=1, lane_ops=46, mufu=2),
    "c1x": dict(desc="BASELINE configs[0] at throughput size (SURVEY.md 8d): the same 3x GBM shape, 1e8 trials x 252 days",
                trials=100_000_000, years=1, lane_ops=46, mufu=2),
    "c2": dict(desc="BASELINE configs[1]: TQQQ/UPRO/SSO, expense ratio + borrow cost + AR(1) Student-t tracking error, "
                    "correlated normal SPY/QQQ shocks, buy&hold, 100k trials x 7560 days",
               trials=100_000, years=30, lane_ops=470, mufu=14),
    "c3": dict(desc="BASELINE configs[2]: regime-switching multivariate-t GARCH/DCC QQQ/SPY/TLT, VIX/IRX, stress + jumps, "
                    "antithetic, moment matching, random start; 1M trials x 7560 days",
               trials=1_000_000, years=30, lane_ops=520 * (7560 + 1260) / 7560, mufu=22),
    "c4": dict(desc="BASELINE configs[3]: C3 market + TQQQ/TMF/SPY layers, S1 S2 S4 S5 S6 S19, per GPU 1.25M trials x 7560 days",
               trials=1_250_000, years=30, lane_ops=1130, mufu=40),
    "c5": dict(desc="BASELINE configs[4]: horizon/leverage grid, L = 3 / 2 / 1 (UPRO / SSO / SPY) on one SPY-like underlying, horizons "
                    "1 2 5 10 20 30 40 y, 4.762M trials per horizon and GPU, per-cell wealth / CAGR / max-drawdown histograms + moments "
                    "reduced on device and allreduced",
               trials=4_762_000, years=None, lane_ops=400, mufu=12, horizons=[1, 2, 5, 10, 20, 30, 40]),
    "boot": dict(desc="legacy_hybrid engine mode (SURVEY.md 8f rank 4): two-pool block bootstrap over a 12 000-row synthetic history "
                      "blended twice with correlated Student-t noise (bootstrap.py:310-458, engine.py:50-80), VIX/IRX, random start; "
                      "TQQQ / SPY / SSO buy&hold, 200k trials x 7560 days per GPU",
                 trials=200_000, years=30, lane_ops=(6 * 110 + 2 * 36 + 40) * (7560 + 1260) / 7560 + 3 * 150, mufu=40),
    "series": dict(desc="materialised-path mode (run_strategy_fixed seam, strategy.py:55): strategy rules S1 S2 S4 S5 on caller-supplied "
                        "day-major fp64 TQQQ_Ret / SPY_Ret series + int32 regime path resident in HBM, 200k paths x 2520 days",
                   trials=200_000, years=10, lane_ops=120, mufu=2, bytes_per_path_day=2 * 8 + 4),
    "series19": dict(desc="materialised-path mode with the reference's default 19-strategy list (typed-slot series build): "
                          "caller-supplied day-major fp64 TQQQ / SSO / TMF / SPY return series + VIX + int32 regime path resident "
                          "in HBM, 100k paths x 2520 days",
                     trials=100_000, years=10, lane_ops=1500, mufu=40, bytes_per_path_day=5 * 8 + 4),
    "sub5": dict(desc="a strategy list with no typed build (generic split pipeline: uncompacted pair producer -> six-field records -> "
                      "dynamic-slot rule kernel): S1 S4 S7 S10 S15 on the institutional market, 227 328 trials x 7560 days per GPU",
                 trials=227_328, years=30, lane_ops=520 * (7560 + 1260) / 7560 + 2 * 150 + 5 * 90 + 200, mufu=40),
    "sma2": dict(desc="SMA switch against buy&hold (S4 on TQQQ, S2 = SPY): a subset of the config-4 list -> its routed six-slot rule "
                      "kernel; 227 328 trials x 7560 days per GPU",
                 trials=227_328, years=30, lane_ops=520 * (7560 + 1260) / 7560 + 2 * 150 + 90 + 60, mufu=30),
    "sweep19": dict(desc="the reference's default sweep: all 19 strategies S1..S19 on the institutional market (config.py:348-437), "
                         "227 328 trials x 7560 days per GPU (whole waves of both kernels of the split pipeline: 2 x 56 832 "
                         "antithetic pairs, 6 x 37 888 rule-kernel trials; at 200 000 trials both run 88 % full)",
                    trials=227_328, years=30, lane_ops=2600, mufu=60),
}


def build_workload(name):
    from letf_monte_carlo_b200 import synthetic
    from letf_monte_carlo_b200.params import build_params
    if name in ("c1", "c1x"):
        ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                        USE_ANTITHETIC_VARIATES=False, LETF_CONFIGS=synthetic.pure_leverage_assets())
        ns.ASSETS = ns.LETF_CONFIGS
        sids = ["B_UPRO"]
        model = synthetic.normal_regime_model("c1")
    elif name == "c2":
        ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                        USE_ANTITHETIC_VARIATES=False)
        sids = ["S1", "B_UPRO", "S3"]
        model = synthetic.normal_regime_model("c2")
    elif name == "c3":
        # LETF_BENCH_NO_MM=1: moment matching off (SURVEY.md 8d asks for both; same pair kernel, zero shift)
        ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS,
                                        USE_MOMENT_MATCHING=os.environ.get("LETF_BENCH_NO_MM", "0") in ("", "0"))
        sids = ["S2", "B_QQQ", "B_TLT"]
        model = synthetic.fallback_regime_model(os.environ.get("LETF_BENCH_MODEL", "default"))
    elif name == "c4":
        ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS)
        sids = ["S1", "S2", "S4", "S5", "S6", "S19"]
        model = synthetic.fallback_regime_model(os.environ.get("LETF_BENCH_MODEL", "default"))
    elif name == "c5":
        ns = synthetic.config_namespace(EXTRA_STRATEGIES=synthetic.EXTRA_BENCHMARKS, USE_RANDOM_START=False,
                                        USE_ANTITHETIC_VARIATES=False)
        sids = ["B_UPRO", "S3", "S2"]
        model = synthetic.normal_regime_model("c2")
        params, keep = build_params(sids, 40, model, ns)
        return ns, model, sids, params, keep
    elif name == "boot":
        ns = synthetic.config_namespace(SIM_ENGINE_MODE="legacy_hybrid")
        sids = ["S1", "S2", "S3"]
        model = synthetic.legacy_bootstrap_model()
    elif name == "sma2":
        ns = synthetic.config_namespace()
        sids = ["S4", "S2"]
        model = synthetic.fallback_regime_model(os.environ.get("LETF_BENCH_MODEL", "default"))
    elif name == "sub5":
        ns = synthetic.config_namespace()
        sids = ["S1", "S4", "S7", "S10", "S15"]
        model = synthetic.fallback_regime_model(os.environ.get("LETF_BENCH_MODEL", "default"))
    elif name == "sweep19":
        ns = synthetic.config_namespace()
        sids = ["S%d" % i for i in range(1, 20)]
        model = synthetic.fallback_regime_model(os.environ.get("LETF_BENCH_MODEL", "default"))
    else:
        raise KeyError(name)
    nu = os.environ.get("LETF_BENCH_JOINT_NU")        # e.g. "5.3,4.2": calibrated models have non-integer dof (calibration.py:399)
    if nu and model.get("joint_return_model"):
        for r, v in enumerate(nu.split(",")):
            model["joint_return_model"]["regimes"][r]["nu"] = float(v)
    params, keep = build_params(sids, WORKLOADS[name]["years"], model, ns)
    return ns, model, sids, params, keep


def out_bytes_fn(n, S):
    return n * (S * (8 + 8 + 4) + 4 + 4 + 8 + 4)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "sm_max_mhz": d["sm_max_mhz"], "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_port_rate(params, budget_s, threads, seed=1):
    """path-days/s of the CPU port (oracle) incl. its own random draws, `threads` host threads,
    on as many trials as fit in ~budget_s seconds.  Returns (rate, n_trials, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.lib()
    T = params.sim_days
    chunk = 4

    def work(i):
        tp = orc.fill_tape(params, chunk, seed * 100003 + i, max_spells=512)
        orc.run(params, tp)
        return chunk * (2 if params.antithetic_pairs else 1)

    t0 = time.perf_counter(); k = work(0); dt1 = time.perf_counter() - t0          # calibration chunk
    n_chunks = max(threads, int(budget_s / max(dt1, 1e-4)) * threads)
    n_chunks = min(n_chunks, 4096)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        done = sum(ex.map(work, range(1, n_chunks + 1)))
    dt = time.perf_counter() - t0
    return done * T / dt, done, dt


def numpy_reference_leg(workload, years, cores):
    """The reference's own NumPy path (unmodified package under baseline/_ref), timed on this box's host cores with the
    reference's default worker count (cpu_count - 2, config.py:151) and with every core; SURVEY.md 8(d)."""
    from baseline import numpy_reference as nr
    if not nr.available():
        return {"unavailable": "baseline/_ref/letf not present"}
    if workload in ("c1", "c1x"):
        d = nr.numpy_rate("c1", 1, 0, 1)
        if "unavailable" in d:
            return d
        return {"value": d["value"], "unit": UNIT, "cores": 1, "kind": "reference", "entry": d["entry"],
                "sample": f"{d['trials']} trials x 252 days in {d['seconds']:.2f} s (vectorised NumPy, one core)"}
    trials = int(os.environ.get("LETF_BENCH_NUMPY_TRIALS", "0")) or 2 * cores
    runs = {}
    for label, w in (("workers_default", max(1, cores - 2)), ("workers_all", cores)):
        d = nr.numpy_rate(workload, years, trials, w)
        if "unavailable" in d:
            return d
        runs[label] = {"workers": w, "value": d["value"], "seconds": d["seconds"], "trials": d["trials"]}
    best = max(runs.values(), key=lambda r: r["value"])
    return {"value": best["value"], "unit": UNIT, "cores": cores, "kind": "reference", "runs": runs, "entry": d["entry"],
            "strategies": d.get("strategies"),
            "sample": f"{trials} trials x {int(years * 252)} days through the unmodified reference (baseline/_ref), full "
                      f"institutional_v1 engine with all six ETF layers (it cannot switch them off), one warm-up call excluded"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns, model, sids, params, keep = build_workload(args.workload)
    threads = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    per_step_budget = 2.0
    for _ in range(args.warmup):
        cpu_port_rate(params, 0.2, threads)
    rates, n_done, secs = [], 0, 0.0
    for s in range(args.steps):
        r, n, dt = cpu_port_rate(params, per_step_budget, threads, seed=s + 2)
        rates.append(r); n_done += n; secs += dt
    value = n_done * params.sim_days / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "trials_per_step_sample": n_done // max(args.steps, 1),
                   "sim_days": params.sim_days},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_done} trials x {params.sim_days} days of the same workload over {args.steps} steps, "
                                   f"oracle/letf_oracle.c (plain C restatement of the reference's NumPy path, xoshiro + "
                                   f"polar/Marsaglia-Tsang draws), {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from letf_monte_carlo_b200 import abi
    from letf_monte_carlo_b200.distributed import exchange_summary
    from letf_monte_carlo_b200.engine import PathEngine
    from letf_monte_carlo_b200.mc_runner import parallel_monte_carlo_fixed

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print meanwhile (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if ws != args.gpus:
        if ws == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ns, model, sids, params, keep = build_workload(args.workload)
    wl = WORKLOADS[args.workload]
    n = int(args.trials or wl["trials"])          # per GPU (weak scaling)
    T = params.sim_days
    first = rank * n
    eng = PathEngine(params, keep, precision=args.precision, device=local)
    out = None
    sm = eng.new_summary(bench_strategy=-1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(ev=None):
        nonlocal out
        flush.zero_()
        if ev:
            ev[0].record()
        out = eng.run(n, first_trial=first, seed=args.seed, out=out)
        if ev:
            ev[1].record()
        eng.reset_summary(sm)
        eng.reduce(out, sm)
        if ws > 1:
            exchange_summary(sm, engine=eng)

    for _ in range(args.warmup):
        step()
    barrier()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
        time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launches
    barrier()
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_end.record()
    barrier()
    launches = eng.launches - launches0
    total_ms = t_start.elapsed_time(t_end)
    # exchange-step correctness (every rank): after the allreduce each strategy's histogram holds ws * n trials minus the
    # ones the reduction rejected as non-finite, and the valid counters agree with it
    hist_n = sm.hist.sum(dim=2).cpu().numpy()
    cnt = sm.counters.cpu().numpy()
    assert (hist_n == cnt[:, :1]).all(), "histogram totals disagree with the valid-trial counters"
    assert (cnt[:, 0] <= ws * n).all() and (cnt[:, 0] >= ws * n - cnt[:, 3]).all() and cnt[:, 0].min() > 0.99 * ws * n, \
        f"summary exchange lost trials: valid {cnt[:, 0]} of {ws * n}"
    summary_check = {"valid": int(cnt[:, 0].min()), "expected": ws * n}
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    clocks = clk.stop() if rank == 0 else None
    info = eng.launch_info()

    # ---- end to end through the public API (the call a user makes): parameter dicts in from the host, the run's summary
    # table out on the host.  summary_only=True: the per-trial arrays stay on the device, the reduction kernel and the
    # cross-rank exchange run inside the call, and only the packed summary (histograms, moments, counters) is copied back.
    # The per-trial route (host result arrays, what the reference returns) is timed beside it.
    h2d = ctypes.sizeof(abi.Params)
    S = len(sids)
    ns.NUM_SIMULATIONS = n * ws
    ns.GPU_SEED = args.seed
    e2e_steps = max(1, min(args.steps, 5))
    kw = dict(cfg=ns, precision=args.precision, device=local, verbose=False)
    for _ in range(max(args.warmup, 3)):          # untimed warm-up calls (engine creation, scratch and result-buffer allocation)
        res = parallel_monte_carlo_fixed(sids, wl["years"], model, {}, summary_only=True, **kw)
        res.table()
    d2h = res.d2h_bytes
    barrier()
    t0 = time.perf_counter()
    call_ms = []
    for _ in range(e2e_steps):
        tc = time.perf_counter()
        res = parallel_monte_carlo_fixed(sids, wl["years"], model, {}, summary_only=True, **kw)
        rows = res.table()                                           # the step's result is read on the host
        med = float(rows[0]["median"])
        call_ms.append(round(1e3 * (time.perf_counter() - tc), 3))
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_launches = 2 * e2e_steps + (e2e_steps if ws > 1 else 0)
    # per-trial route: every trial's results copied to host arrays (bounded: at most 2e7 trials per rank)
    pt = None
    if n <= 20_000_000:
        d2h_pt = n * (S * (8 + 8 + 4) + 4 + 4 + 8 + 4)
        for _ in range(2):
            r2 = parallel_monte_carlo_fixed(sids, wl["years"], model, {}, **kw)
        barrier()
        t1 = time.perf_counter()
        pt_steps = max(1, min(args.steps, 2))
        for _ in range(pt_steps):
            r2 = parallel_monte_carlo_fixed(sids, wl["years"], model, {}, **kw)
            mean_w = float(np.mean(r2[sids[0]].final_wealth))
        barrier()
        pt = {"value": ws * n * T * pt_steps / (time.perf_counter() - t1), "unit": UNIT, "d2h_bytes_per_step": d2h_pt,
              "steps": pt_steps, "mean_final_wealth": mean_w,
              "api": "parallel_monte_carlo_fixed(...): host result arrays out (library's pinned staging buffer)"}
        e2e_launches += pt_steps

    t = torch.tensor([total_ms, kern_ms, e2e_s], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_s = (float(x) for x in t.cpu())

    if rank == 0:
        pk = peaks()
        ms_per_step = total_ms / args.steps
        value = ws * n * T / (ms_per_step * 1e-3)
        sm_count = info["sm_count"]
        peak_lane = sm_count * 128 * pk["sm_max_mhz"] * 1e6 / 1e12               # T lane-ops/s at max clock
        achieved = n * T * wl["lane_ops"] / (kern_ms * 1e-3) / 1e12
        out_bytes = out_bytes_fn(n, S)
        ncu_c = {}
        try:
            ncu_c = json.load(open(os.path.join(ROOT, "profiles", "ncu_constants.json"))).get(args.workload, {})
        except Exception:
            pass
        traffic = ncu_c["dram_bytes_per_trial"] * n if "dram_bytes_per_trial" in ncu_c else None
        exe = ncu_c.get("executed_warp_instr_per_path_day")
        # algorithmic HBM bytes of the launch: per-trial outputs, plus - for the two-pass (moment matching) kernels - one
        # 32-byte record per generated day and antithetic pair written by pass 1 and read back by pass 2
        G = T + int(params.buffer_days)
        alg_bytes = out_bytes_fn(n, S)
        if info["kernel"].startswith("f32_joint_t_pair") or "prod" in info["kernel"]:
            alg_bytes += n * G * 32
        if "+rules_" in info["kernel"]:      # split pipeline: the per-day records are written once and read once (4 B per field)
            nf = 5 if ("rules_c4" in info["kernel"] or "rules_sub6" in info["kernel"]) else (8 if "rules_all" in info["kernel"] else 6)
            alg_bytes += n * T * nf * 4 * 2
        pipes = None
        try:
            pp = json.load(open(os.path.join(ROOT, "profiles", "pipe_peaks.json")))
            pipes = {k: v["lane_ops_per_clk_per_sm"] for k, v in pp.items() if isinstance(v, dict)}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "trials_per_gpu": n, "sim_days": T,
                       "strategies": sids, "parallelism": f"trials sharded over {ws} GPU(s), no data-path collective; "
                                                          f"summary allreduce (NCCL) per step",
                       "rng": "on-device Philox4x32-10 keyed by (seed, sim_id, day, stream)",
                       "l2": "256 MiB memset between steps (inside the timed region); the path kernel reads no HBM inputs",
                       "kernel": {"name": info["kernel"], "grid": info["grid"], "block": info["block"], "regs": info["regs_per_thread"],
                                  "scratch_bytes": info["scratch_bytes"]}},
            "clocks": clocks,
            "e2e": {"value": ws * n * T * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "letf_monte_carlo_b200.parallel_monte_carlo_fixed(..., summary_only=True).table(): parameter dicts in, "
                           "summary table out on the host (reduction + cross-rank exchange inside the call; engine handle "
                           "cached across calls)",
                    "steps": e2e_steps, "median_final_wealth_top_row": med, "rank0_call_ms": call_ms, "per_trial_route": pt},
            "gpu_launches": launches + e2e_launches,
            "summary_check": summary_check,
            "roofline": {"bound": "alu", "achieved": achieved, "peak": peak_lane, "unit": "T lane-ops/s", "frac": achieved / peak_lane,
                         "traffic": traffic, "algorithmic_bytes": alg_bytes, "kernel": info["kernel"], "kernel_ms": kern_ms,
                         "pipe_peaks_lane_ops_per_clk_per_sm": pipes,
                         "issue_slot_frac_executed": (n * T * exe / (kern_ms * 1e-3) / 1e12 / peak_lane) if exe else None,
                         "ncu": ncu_c.get("source"),
                         "lane_ops_per_path_day": wl["lane_ops"], "units_per_launch": n * T,
                         "peak_def": f"{sm_count} SMs x 128 FP32/INT lanes x {pk['sm_max_mhz']:.0f} MHz ({pk['source']})",
                         "frac_at_observed_clock": (achieved / (sm_count * 128 * clocks["sm_mhz"] * 1e6 / 1e12)
                                                    if clocks and clocks.get("sm_mhz") else None),
                         "hbm": {"achieved_gbs": (traffic if traffic is not None else out_bytes) / (kern_ms * 1e-3) / 1e9,
                                 "peak_gbs": pk["hbm_gbs"],
                                 "note": ("DRAM bytes of the committed ncu capture scaled to this launch (two-pass scratch, history "
                                          "rings, outputs) over the live kernel time" if traffic is not None else
                                          "outputs only; per-day paths never round-trip HBM")}},
        }
        if ws == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate, nt, dt = cpu_port_rate(params, 1.5, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{nt} trials x {T} days of the same workload in {dt:.1f} s, oracle/letf_oracle.c "
                                              f"incl. its own CPU random draws, {threads} threads"}
            line["cpu_baseline_numpy"] = numpy_reference_leg(args.workload, wl["years"], threads)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    eng.close()
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(args):
    """BASELINE configs[4]: a step = one evaluation of the whole horizon/leverage grid (7 launches, 7 on-device
    reductions, per-cell summary allreduce); only histograms / moments / counters ever leave the device."""
    import torch
    import torch.distributed as dist
    from letf_monte_carlo_b200.grid import HorizonGrid

    rank = int(os.environ.get("RANK", "0")); ws = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ns, model, sids, params, keep = build_workload("c5")
    wl = WORKLOADS["c5"]
    n = int(args.trials or wl["trials"])
    grid = HorizonGrid(wl["horizons"], sids, model, ns, precision=args.precision, device=local)
    path_days = grid.path_days(n)                      # per GPU and step

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        grid.run(n * ws, seed=args.seed)
    barrier()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start(); time.sleep(0.25)
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = grid.launches
    barrier()
    t0e.record()
    for _ in range(args.steps):
        grid.run(n * ws, seed=args.seed)
    t1e.record()
    barrier()
    total_ms = t0e.elapsed_time(t1e)
    launches = grid.launches - l0
    clocks = clk.stop() if rank == 0 else None
    # end to end: the grid evaluated and its summary table read on the host
    barrier()
    t0 = time.perf_counter()
    rows = grid.run(n * ws, seed=args.seed + 1).table()
    barrier()
    e2e_s = time.perf_counter() - t0
    d2h = sum(int(sm.hist.numel() * 8 + sm.counters.numel() * 8 + sm.moments.numel() * 8) for sm in grid.summaries)
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = (float(x) for x in t.cpu())
    if rank == 0:
        pk = peaks()
        ms = total_ms / args.steps
        peak_lane = 148 * 128 * pk["sm_max_mhz"] * 1e6 / 1e12
        achieved = path_days * wl["lane_ops"] / (ms * 1e-3) / 1e12
        med30 = [r for r in rows if r["horizon"] == 30.0]
        line = {"metric": METRIC, "value": ws * path_days / (ms * 1e-3), "unit": UNIT, "n_gpus": ws, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": f"c5: {wl['desc']}", "trials_per_horizon_per_gpu": n, "horizons_years": wl["horizons"],
                           "strategies": sids, "path_days_per_gpu_step": path_days,
                           "l2": "outputs of each launch (>= 95 MB) exceed nothing the next launch reads; the path kernels read no HBM inputs"},
                "clocks": clocks,
                "e2e": {"value": ws * path_days / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 7 * ctypes.sizeof(type(params)),
                        "d2h_bytes_per_step": d2h, "api": "letf_monte_carlo_b200.grid.HorizonGrid.run(...).table()",
                        "median_wealth_30y": {r["id"]: r["p50"] for r in med30}},
                "gpu_launches": launches + 14,
                "roofline": {"bound": "alu", "achieved": achieved, "peak": peak_lane, "unit": "T lane-ops/s", "frac": achieved / peak_lane,
                             "traffic": None, "kernel": "letf_path_kernel (7 launches) + letf_reduce_kernel",
                             "lane_ops_per_path_day": wl["lane_ops"], "units_per_launch": path_days,
                             "note": "whole-step time (reductions included), not a single launch"}}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    grid.close()
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_series(args):
    """The HBM-side mode: the strategy kernel streams materialised return series (letf_run_strategies).  Roofline bound
    = HBM: algorithmic bytes = the series it must read once (2 x fp64 + the int32 regime per path-day)."""
    import torch
    from letf_monte_carlo_b200 import synthetic
    from letf_monte_carlo_b200.engine import PathEngine
    from letf_monte_carlo_b200.params import build_params
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("the series workload is a single-GPU measurement")
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    wl = WORKLOADS[args.workload]
    n, T = int(args.trials or wl["trials"]), int(wl["years"] * 252)
    ns = synthetic.config_namespace()
    full = args.workload == "series19"
    sids = ["S%d" % i for i in range(1, 20)] if full else ["S1", "S2", "S4", "S5"]
    params, keep = build_params(sids, wl["years"], synthetic.fallback_regime_model(), ns)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(args.seed)
    spy = (0.0003 + 0.01 * torch.randn((T, n), device=dev, dtype=torch.float64, generator=g))
    tqqq = (3.3 * spy - 0.0003)
    regime = (torch.rand((T, n), device=dev, generator=g) < 0.2).to(torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    eng = PathEngine(params, keep, precision=args.precision, device=0)
    series = {"TQQQ": tqqq, "SPY": spy}
    vix = None
    if full:
        series["SSO"] = 2.0 * spy - 0.0001
        series["TMF"] = -0.6 * spy + 0.004 * torch.randn((T, n), device=dev, dtype=torch.float64, generator=g)
        vix = 18.0 + 6.0 * torch.rand((T, n), device=dev, dtype=torch.float64, generator=g) + 14.0 * regime.to(torch.float64)

    def step():
        flush.zero_()
        return eng.run_strategies(series, regime=regime, vix=vix)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    clk = ClockSampler(0); clk.start(); time.sleep(0.25)
    kms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        kms.append(eng.launch_info()["last_kernel_ms"])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = clk.stop()
    kern_ms = float(np.mean(kms))
    info = eng.launch_info()
    pk = peaks()
    bytes_launch = n * T * wl["bytes_per_path_day"]
    achieved = bytes_launch / (kern_ms * 1e-3) / 1e9
    # end to end: host series in (pinned), results out
    hs = {k: v.cpu().pin_memory() for k, v in series.items()}
    hr = regime.cpu().pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hv = vix.cpu().pin_memory() if vix is not None else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = eng.run_strategies({k: v.to(dev, non_blocking=True) for k, v in hs.items()}, regime=hr.to(dev, non_blocking=True),
                             vix=hv.to(dev, non_blocking=True) if hv is not None else None)
    fw = out.to_host().final_wealth
    e2e_s = time.perf_counter() - t0
    line = {"metric": METRIC, "value": n * T * args.steps / wall, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "paths": n, "sim_days": T, "strategies": sids,
                       "l2": f"256 MiB memset between steps; inputs ({bytes_launch / 1e9:.1f} GB) exceed L2",
                       "kernel": {"name": info["kernel"], "grid": info["grid"], "block": info["block"], "regs": info["regs_per_thread"],
                                  "launches_per_step": int(info["kernel_launches"])}},
            "clocks": clocks,
            "e2e": {"value": n * T / e2e_s, "unit": UNIT, "h2d_bytes_per_step": bytes_launch, "d2h_bytes_per_step": n * len(sids) * 20,
                    "api": "PathEngine.run_strategies with pinned host series", "mean_final_wealth": float(np.mean(fw[0]))},
            "gpu_launches": (args.steps + 1) * max(1, int(info["kernel_launches"])),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": None, "kernel": info["kernel"] or "letf_path_kernel<series>", "kernel_ms": kern_ms,
                         "bytes_per_path_day": wl["bytes_per_path_day"], "units_per_launch": n * T,
                         "peak_def": f"measured copy bandwidth ({pk['source']})"}}
    eng.close()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--trials", type=int, default=0, help="trials per GPU (default: the workload's)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--seed", type=int, default=20240229)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    elif args.workload in ("series", "series19"):
        run_series(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
