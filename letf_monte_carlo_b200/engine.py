"""`PathEngine`: Python handle of the CUDA path engine (thin wrapper of the C-ABI).

PyTorch is used only for device memory, the current stream and (in mc_runner / bench)
`torch.distributed`; every result is produced by the kernels in `libletf_b200.so`.
There is no CPU fallback: constructing a PathEngine without the built library or
without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .outputs import DeviceOutputs, HostOutputs
from .tape import HostTape

_PREC = {"f64": abi.PREC_F64, "f32": abi.PREC_F32, abi.PREC_F64: abi.PREC_F64, abi.PREC_F32: abi.PREC_F32}


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"letf engine error {code}: {msg}")
        self.code = code


class NonFiniteTrials(ValueError):
    """Some trial produced non-finite returns or VIX <= 0; the reference raises ValueError
    out of the trial (engine.py:486-498)."""


def _trade_dicts(day, asset, action, shares, price, n: int):
    """the reference's Trade_List (trade.py:83-91) over the replay buffers: a `TradeList` (reads like the list of dicts, keeps
    the columns)"""
    from .trade import TradeList
    return TradeList(day[:n].copy(), asset[:n].copy(), action[:n].copy(), shares[:n].copy(), price[:n].copy(), abi.ASSET_NAMES)


def default_device() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


class DeviceTape:
    """A HostTape uploaded to the device (torch tensors keep the memory alive)."""

    def __init__(self, host: HostTape, device):
        import torch
        self.host = host
        self.t = {k: torch.from_numpy(np.ascontiguousarray(a)).to(device) for k, a in host.arrays().items()}

    def as_struct(self) -> abi.Tape:
        s = abi.Tape()
        h = self.host
        s.n_tapes, s.gen_days, s.sim_days, s.max_spells = h.n_tapes, h.gen_days, h.sim_days, h.max_spells
        s.max_blocks = h.max_blocks
        s.boot_block_size = h.boot_block_size
        for k, t in self.t.items():
            setattr(s, k, t.data_ptr() if t.numel() else None)
        return s

    def to_host(self) -> HostTape:
        for k, t in self.t.items():
            getattr(self.host, k)[...] = t.cpu().numpy()
        return self.host


class DeviceSummary:
    """Histogram / moment buffers of `letf_reduce` (sum/min/max-reducible across GPUs).

    One packed buffer of 8-byte words, [hist S*3*4096 | counters S*4 | moments S*3*2 | extrema S*3*2]
    (include/letf_b200.h `letf_summary_words`); `hist`, `counters`, `moments`, `extrema` are views into it, so the
    ranks of a multi-GPU run exchange the whole summary with one all-gather (distributed.exchange_summary)."""

    def __init__(self, n_strategies: int, device, lo=(-20.0, -1.0, 0.0), hi=(10.0, 2.0, 1.0), bench_strategy=-1):
        import torch
        S = max(int(n_strategies), 1)
        self.n_strategies = int(n_strategies)
        nh, nc, nm = S * abi.N_METRICS * abi.HIST_BINS, S * 4, S * abi.N_METRICS * 2
        self.buf = torch.zeros(nh + nc + 2 * nm, dtype=torch.int64, device=device)
        self.hist = self.buf[:nh].view(S, abi.N_METRICS, abi.HIST_BINS)
        self.counters = self.buf[nh:nh + nc].view(S, 4)
        self.moments = self.buf[nh + nc:nh + nc + nm].view(torch.float64).view(S, abi.N_METRICS, 2)
        self.extrema = self.buf[nh + nc + nm:].view(torch.float64).view(S, abi.N_METRICS, 2)
        self.extrema[:, :, 0] = float("inf")
        self.extrema[:, :, 1] = float("-inf")
        self.lo, self.hi = tuple(map(float, lo)), tuple(map(float, hi))
        self.bench_strategy = int(bench_strategy)

    def as_struct(self) -> abi.Summary:
        s = abi.Summary()
        s.hist, s.moments = self.hist.data_ptr(), self.moments.data_ptr()
        s.extrema, s.counters = self.extrema.data_ptr(), self.counters.data_ptr()
        for q in range(abi.N_METRICS):
            s.hist_lo[q], s.hist_hi[q] = self.lo[q], self.hi[q]
        s.bench_strategy = self.bench_strategy
        return s

    def bin_edges(self, metric: int) -> np.ndarray:
        return np.linspace(self.lo[metric], self.hi[metric], abi.HIST_BINS + 1)


class PathEngine:
    def __init__(self, params: abi.Params, keep=(), precision="f32", device: int | None = None):
        import torch
        self._lib = abi.load_engine_library()
        if not torch.cuda.is_available():
            raise RuntimeError("PathEngine needs a CUDA device (B200 / sm_100a); there is no CPU path")
        self.device_index = default_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.params = params
        self._keep = list(keep)
        self.precision = _PREC[precision]
        h = C.c_void_p()
        rc = self._lib.letf_create(C.byref(params), self.device_index, self.precision, C.byref(h))
        self._check(rc)
        self._h = h
        self.launches = 0          # kernels launched through this handle (bench.py reports it)

    # ------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc == abi.OK:
            return
        msg = (self._lib.letf_last_error() or b"").decode()
        if rc == abi.ERR_NONFINITE:
            raise NonFiniteTrials(msg)
        raise EngineError(rc, msg)

    @staticmethod
    def raise_if_nonfinite(flags):
        """the device-buffer route's counterpart of letf_run_host's check: the reference raises ValueError out of a trial
        with non-finite returns or VIX <= 0 (engine.py:486-498)"""
        bad = int(np.count_nonzero(np.asarray(flags) & (abi.FLAG_NONFINITE_RET | abi.FLAG_BAD_VIX)))
        if bad:
            raise NonFiniteTrials(f"{bad} of {len(flags)} trials generated non-finite returns or VIX <= 0 "
                                  f"(reference raises ValueError, engine.py:486-498)")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.letf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_strategies(self):
        return int(self.params.n_strategies)

    def _stream(self, stream):
        import torch
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return C.c_void_p(s.cuda_stream)

    # ------------------------------------------------------------------------------------
    def run(self, n_trials: int, *, first_trial: int = 0, seed: int = 0, rng_mode: int = abi.RNG_PHILOX,
            tape: DeviceTape | None = None, out: DeviceOutputs | None = None, journal: bool = False,
            stream=None) -> DeviceOutputs:
        """Simulate trials [first_trial, first_trial + n_trials) into device buffers (asynchronous)."""
        if out is None:
            out = DeviceOutputs(self.n_strategies, n_trials, self.device, journal=journal)
        ts = tape.as_struct() if tape is not None else None
        rc = self._lib.letf_run(self._h, self._stream(stream), seed, first_trial, n_trials, rng_mode,
                                C.byref(ts) if ts is not None else None, C.byref(out.as_struct()))
        self._check(rc)
        self.launches += 1 if n_trials > 0 else 0
        return out

    def run_host(self, n_trials: int, *, first_trial: int = 0, seed: int = 0, rng_mode: int = abi.RNG_PHILOX,
                 tape: HostTape | None = None, out: HostOutputs | None = None, journal: bool = False,
                 raise_nonfinite: bool = True, pinned: bool = False) -> HostOutputs:
        """Same with host (numpy) buffers: copies in, runs, copies out, synchronises."""
        if out is None:
            out = HostOutputs(self.n_strategies, n_trials, journal=journal, pinned=pinned)
        ts = tape.as_struct() if tape is not None else None
        rc = self._lib.letf_run_host(self._h, seed, first_trial, n_trials, rng_mode,
                                     C.byref(ts) if ts is not None else None, C.byref(out.as_struct()))
        self.launches += 1 if n_trials > 0 else 0
        if rc == abi.ERR_NONFINITE and not raise_nonfinite:
            return out
        self._check(rc)
        return out

    def run_strategies(self, asset_ret: dict, *, tlt_ret=None, vix=None, regime=None, cash_ret=None, journal: bool = False,
                       stream=None) -> DeviceOutputs:
        """Strategy rules on caller-supplied series (`run_strategy_fixed` seam, strategy.py:55).
        asset_ret: {asset name: [T][n] array-like}; all arrays day-major."""
        import torch
        dev = self.device

        def up(a, dt):
            if a is None:
                return None
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
            return t.to(device=dev, dtype=dt).contiguous()

        tens = {k: up(v, torch.float64) for k, v in asset_ret.items()}
        first = next(iter(tens.values()))
        T, n = int(first.shape[0]), int(first.shape[1])
        s = abi.Series()
        s.n_paths, s.sim_days = n, T
        for name, t in tens.items():
            assert tuple(t.shape) == (T, n)
            s.asset_ret[abi.ASSET_INDEX[name]] = t.data_ptr()
        tl, vx, rg = up(tlt_ret, torch.float64), up(vix, torch.float64), up(regime, torch.int32)
        cr = up(cash_ret, torch.float64)
        s.cash_ret = cr.data_ptr() if cr is not None else None
        s.tlt_ret = tl.data_ptr() if tl is not None else None
        s.vix = vx.data_ptr() if vx is not None else None
        s.regime = rg.data_ptr() if rg is not None else None
        out = DeviceOutputs(self.n_strategies, n, dev, journal=journal)
        rc = self._lib.letf_run_strategies(self._h, self._stream(stream), C.byref(s), C.byref(out.as_struct()))
        self._check(rc)
        self.launches += 1 if n > 0 else 0
        torch.cuda.current_stream(dev).synchronize()   # inputs above must outlive the kernel
        return out

    def new_summary(self, **kw) -> DeviceSummary:
        return DeviceSummary(self.n_strategies, self.device, **kw)

    def reduce(self, out: DeviceOutputs, summary: DeviceSummary, n_trials: int | None = None, stream=None):
        n = out.n if n_trials is None else n_trials
        rc = self._lib.letf_reduce(self._h, self._stream(stream), C.byref(out.as_struct()), n,
                                   C.byref(summary.as_struct()))
        self._check(rc)
        self.launches += 1 if n > 0 else 0
        return summary

    def combine_summaries(self, gathered, world: int, summary: DeviceSummary, stream=None):
        """fold `world` gathered packed summaries ([world][words] int64 device tensor) into `summary` (one kernel)"""
        assert gathered.is_cuda and gathered.is_contiguous() and gathered.numel() == world * summary.buf.numel()
        rc = self._lib.letf_summary_combine(self._h, self._stream(stream), gathered.data_ptr(), int(world),
                                            max(self.n_strategies, 1), summary.buf.data_ptr())
        self._check(rc)
        self.launches += 1

    def reset_summary(self, summary: DeviceSummary, stream=None):
        self._check(self._lib.letf_summary_reset(self._h, self._stream(stream), C.byref(summary.as_struct())))
        self.launches += 1

    def trade_list(self, trial: int, strategy: int, *, seed: int = 0, rng_mode: int = abi.RNG_PHILOX,
                   tape: HostTape | None = None, cap: int = 1 << 16) -> list[dict]:
        """`Trade_List` of one trial / one taxable strategy (trade.py:83-91), by replaying the trial."""
        day, asset, action = (np.zeros(cap, np.int32) for _ in range(3))
        shares, price = np.zeros(cap), np.zeros(cap)
        cnt = C.c_int32(0)
        ts = tape.as_struct() if tape is not None else None
        rc = self._lib.letf_trade_list_host(self._h, seed, trial, strategy, rng_mode,
                                            C.byref(ts) if ts is not None else None, cap, day.ctypes.data,
                                            asset.ctypes.data, action.ctypes.data, shares.ctypes.data,
                                            price.ctypes.data, C.byref(cnt))
        self._check(rc)
        self.launches += 1
        if cnt.value > cap:
            return self.trade_list(trial, strategy, seed=seed, rng_mode=rng_mode, tape=tape, cap=cnt.value)
        return _trade_dicts(day, asset, action, shares, price, cnt.value)

    def trade_lists(self, trials, strategy: int, *, seed: int = 0, cap: int = 2048) -> list[list[dict]]:
        """`Trade_List`s of several trials of one taxable strategy in ONE launch (the <= 50 sampled trials of
        reporting.py:457-467); a single-thread replay per trial would cost one launch each."""
        ids = np.ascontiguousarray(np.asarray(list(trials), dtype=np.int64))
        n = len(ids)
        if n == 0:
            return []
        while True:
            day, asset, action = (np.zeros((n, cap), np.int32) for _ in range(3))
            shares, price = np.zeros((n, cap)), np.zeros((n, cap))
            cnt = np.zeros(n, np.int32)
            rc = self._lib.letf_trade_lists_host(self._h, seed, ids.ctypes.data, n, strategy, cap, day.ctypes.data,
                                                 asset.ctypes.data, action.ctypes.data, shares.ctypes.data,
                                                 price.ctypes.data, cnt.ctypes.data)
            self._check(rc)
            self.launches += 1
            if cnt.max(initial=0) <= cap:
                break
            cap = int(cnt.max())
        return [_trade_dicts(day[j], asset[j], action[j], shares[j], price[j], int(cnt[j])) for j in range(n)]

    def trade_lists_multi(self, trials, strategies, *, seed: int = 0, cap: int = 2048) -> list:
        """`Trade_List`s of (trial, strategy) pairs in ONE launch - the whole sampled-trial tax table (<= 50 trials for each
        of the taxable strategies, reporting.py:457-467) at the cost of one replay: a replay is bound by the latency of a single
        thread walking its path, not by the number of threads."""
        ids = np.ascontiguousarray(np.asarray(list(trials), dtype=np.int64))
        st = np.ascontiguousarray(np.asarray(list(strategies), dtype=np.int32))
        n = len(ids)
        if n != len(st):
            raise ValueError("trials and strategies must have the same length")
        if n == 0:
            return []
        while True:
            day, asset, action = (np.zeros((n, cap), np.int32) for _ in range(3))
            shares, price = np.zeros((n, cap)), np.zeros((n, cap))
            cnt = np.zeros(n, np.int32)
            rc = self._lib.letf_trade_lists_multi_host(self._h, seed, ids.ctypes.data, st.ctypes.data, n, cap, day.ctypes.data,
                                                       asset.ctypes.data, action.ctypes.data, shares.ctypes.data,
                                                       price.ctypes.data, cnt.ctypes.data)
            self._check(rc)
            self.launches += 1
            if cnt.max(initial=0) <= cap:
                break
            cap = int(cnt.max())
        return [_trade_dicts(day[j], asset[j], action[j], shares[j], price[j], int(cnt[j])) for j in range(n)]

    def equity_curve(self, trial: int, strategy: int, *, seed: int = 0, rng_mode: int = abi.RNG_PHILOX,
                     tape: HostTape | None = None) -> np.ndarray:
        """daily equity of one trial / one strategy (the Series `run_strategy_fixed` returns), by replaying the trial"""
        curve = np.zeros(max(int(self.params.sim_days), 1))
        ts = tape.as_struct() if tape is not None else None
        rc = self._lib.letf_equity_curve_host(self._h, seed, trial, strategy, rng_mode,
                                              C.byref(ts) if ts is not None else None, curve.ctypes.data)
        self._check(rc)
        self.launches += 1
        return curve[:int(self.params.sim_days)]

    def market_returns(self, trial: int, *, seed: int = 0, rng_mode: int = abi.RNG_PHILOX, tape: HostTape | None = None) -> np.ndarray:
        """[sim_days][3] generated SPY / QQQ / TLT returns of one trial's window (before the jump overlay), by replaying it"""
        T = int(self.params.sim_days)
        ret = np.zeros((max(T, 1), 3))
        ts = tape.as_struct() if tape is not None else None
        rc = self._lib.letf_market_returns_host(self._h, seed, trial, rng_mode, C.byref(ts) if ts is not None else None,
                                                ret.ctypes.data)
        self._check(rc)
        self.launches += 1
        return ret[:T]

    def run_strategy_curve(self, frame: dict, strategy: int, *, tlt_ret=None, vix=None, regime=None, cash_ret=None, cap: int = 4096):
        """`run_strategy_fixed` on ONE caller-supplied frame (strategy.py:55): frame = {asset name: [T] returns}.
        -> (equity curve [T], num_trades, Trade_List rows for a taxable id)"""
        cols = {k: np.ascontiguousarray(np.asarray(v, dtype=np.float64)) for k, v in frame.items()}
        T = len(next(iter(cols.values())))
        s = abi.Series()
        s.n_paths, s.sim_days = 1, T
        for name, a in cols.items():
            assert a.shape == (T,)
            s.asset_ret[abi.ASSET_INDEX[name]] = a.ctypes.data
        tl = None if tlt_ret is None else np.ascontiguousarray(np.asarray(tlt_ret, dtype=np.float64))
        vx = None if vix is None else np.ascontiguousarray(np.asarray(vix, dtype=np.float64))
        rg = None if regime is None else np.ascontiguousarray(np.asarray(regime, dtype=np.int32))
        cr = None if cash_ret is None else np.ascontiguousarray(np.asarray(cash_ret, dtype=np.float64))
        s.cash_ret = cr.ctypes.data if cr is not None else None
        for name, a in (("tlt_ret", tl), ("vix", vx), ("regime", rg), ("cash_ret", cr)):
            if a is not None and a.shape != (T,):
                raise ValueError(f"{name} has shape {a.shape}, the frame has {T} rows (the library reads T elements of every column)")
        if rg is not None and rg.size and (rg.min() < 0 or rg.max() > 1):
            raise ValueError("regime path values must be 0 or 1 (they index the per-regime cash rate and spread cost)")
        s.tlt_ret = tl.ctypes.data if tl is not None else None
        s.vix = vx.ctypes.data if vx is not None else None
        s.regime = rg.ctypes.data if rg is not None else None
        curve = np.zeros(max(T, 1))
        nt, cnt = C.c_int32(0), C.c_int32(0)
        while True:
            day, asset, action = (np.zeros(cap, np.int32) for _ in range(3))
            shares, price = np.zeros(cap), np.zeros(cap)
            rc = self._lib.letf_run_strategy_curve_host(self._h, C.byref(s), strategy, curve.ctypes.data, C.byref(nt), cap,
                                                        day.ctypes.data, asset.ctypes.data, action.ctypes.data,
                                                        shares.ctypes.data, price.ctypes.data, C.byref(cnt))
            self._check(rc)
            self.launches += 1
            if cnt.value <= cap:
                break
            cap = cnt.value
        return curve[:T], int(nt.value), _trade_dicts(day, asset, action, shares, price, cnt.value)

    def launch_info(self) -> dict:
        li = abi.LaunchInfo()
        self._check(self._lib.letf_launch_info(self._h, C.byref(li)))
        d = {k: getattr(li, k) for k, _ in abi.LaunchInfo._fields_ if k != "pad"}
        d["kernel"] = d["kernel"].decode()
        return d
