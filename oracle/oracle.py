"""ctypes front end of oracle/letf_oracle.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from letf_monte_carlo_b200 import abi
from letf_monte_carlo_b200.tape import HostTape
from letf_monte_carlo_b200.outputs import HostOutputs  # noqa: F401  (numpy-backed LetfOutputs)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libletf_oracle.so")
_lib = None


class OracleSeries(C.Structure):
    _fields_ = [("regime_full", C.c_void_p), ("gen", C.c_void_p * 5),
                ("asset_ret", C.c_void_p * abi.N_ASSETS),
                ("liq", C.c_void_p), ("cred", C.c_void_p), ("jump", C.c_void_p), ("equity", C.c_void_p)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "letf_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "letf_b200.h")
    stale = (not os.path.exists(LIB_PATH)
             or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(LIB_PATH))
             or (os.path.exists(hdr) and os.path.getmtime(hdr) > os.path.getmtime(LIB_PATH)))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libletf_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.letf_oracle_run.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Tape), C.c_int64, C.c_int64,
                                      C.POINTER(abi.Outputs)]
        L.letf_oracle_run.restype = C.c_int
        L.letf_oracle_run_series.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Tape), C.c_int64,
                                             C.POINTER(abi.Outputs), C.POINTER(OracleSeries)]
        L.letf_oracle_run_series.restype = C.c_int
        L.letf_oracle_trade_list.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Tape), C.c_int64, C.c_int,
                                             C.POINTER(abi.Outputs), C.c_int] + [C.c_void_p] * 5
        L.letf_oracle_trade_list.restype = C.c_int
        L.letf_oracle_run_strategies.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Series),
                                                 C.POINTER(abi.Outputs)]
        L.letf_oracle_run_strategies.restype = C.c_int
        L.letf_oracle_fill_tape.argtypes = [C.POINTER(abi.Params), C.c_uint64, C.POINTER(abi.Tape)]
        L.letf_oracle_fill_tape.restype = C.c_int
        for nm in ("orc_roll_sum", "orc_roll_mean", "orc_roll_std", "orc_roll_skew"):
            f = getattr(L, nm)
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
            f.restype = None
        _lib = L
    return _lib


def run(params: abi.Params, tape: HostTape, first: int = 0, n: int | None = None) -> HostOutputs:
    n_all = tape.n_tapes * (2 if params.antithetic_pairs else 1)
    n = n_all - first if n is None else n
    out = HostOutputs(params.n_strategies, n)
    rc = lib().letf_oracle_run(C.byref(params), C.byref(tape.as_struct()), first, n, C.byref(out.as_struct()))
    if rc < 0:
        raise RuntimeError(f"letf_oracle_run failed rc={rc}")
    out.n_raised = rc
    return out


def run_series(params: abi.Params, tape: HostTape, trial: int):
    T, G, S = params.sim_days, tape.gen_days, max(params.n_strategies, 1)
    out = HostOutputs(params.n_strategies, 1)
    d = {"regime_full": np.zeros(G, np.int32), "gen": np.zeros((5, G)),
         "asset_ret": np.zeros((abi.N_ASSETS, T)), "liq": np.zeros(T), "cred": np.zeros(T),
         "jump": np.zeros(T), "equity": np.zeros((S, T))}
    s = OracleSeries()
    s.regime_full = d["regime_full"].ctypes.data
    for c in range(5):
        s.gen[c] = d["gen"][c].ctypes.data
    for a in range(abi.N_ASSETS):
        s.asset_ret[a] = d["asset_ret"][a].ctypes.data
    s.liq, s.cred, s.jump, s.equity = (d[k].ctypes.data for k in ("liq", "cred", "jump", "equity"))
    rc = lib().letf_oracle_run_series(C.byref(params), C.byref(tape.as_struct()), trial,
                                      C.byref(out.as_struct()), C.byref(s))
    if rc not in (0, abi.ERR_NONFINITE):
        raise RuntimeError(f"letf_oracle_run_series rc={rc}")
    d["out"] = out
    return d


def trade_list(params: abi.Params, tape: HostTape, trial: int, strategy: int, cap: int = 1 << 16):
    out = HostOutputs(params.n_strategies, 1)
    day, asset, action = (np.zeros(cap, np.int32) for _ in range(3))
    shares, price = np.zeros(cap), np.zeros(cap)
    k = lib().letf_oracle_trade_list(C.byref(params), C.byref(tape.as_struct()), trial, strategy,
                                     C.byref(out.as_struct()), cap, day.ctypes.data, asset.ctypes.data,
                                     action.ctypes.data, shares.ctypes.data, price.ctypes.data)
    if k < 0:
        raise RuntimeError("letf_oracle_trade_list failed")
    return [{"day_index": int(day[i]), "asset": abi.ASSET_NAMES[asset[i]],
             "action": "BUY" if action[i] > 0 else "SELL", "shares": float(shares[i]),
             "price": float(price[i]), "dollar_amount": float(shares[i] * price[i])} for i in range(k)]


def fill_tape(params: abi.Params, n_tapes: int, seed: int, max_spells: int = 2048) -> HostTape:
    from letf_monte_carlo_b200.tape import boot_tape_shape
    G = params.sim_days + params.buffer_days
    tp = HostTape(n_tapes, G, params.sim_days, max_spells, *boot_tape_shape(params))
    ts = tp.as_struct()
    lib().letf_oracle_fill_tape(C.byref(params), seed, C.byref(ts))
    return tp


def _roll(name, x, w, mp):
    x = np.ascontiguousarray(x, dtype=float)
    out = np.empty_like(x)
    getattr(lib(), name)(x.ctypes.data, len(x), int(w), int(mp), out.ctypes.data)
    return out


def roll_sum(x, w, mp): return _roll("orc_roll_sum", x, w, mp)
def roll_mean(x, w, mp): return _roll("orc_roll_mean", x, w, mp)
def roll_std(x, w, mp): return _roll("orc_roll_std", x, w, mp)
def roll_skew(x, w, mp): return _roll("orc_roll_skew", x, w, mp)
