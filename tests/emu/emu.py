"""Loader for the TEST-ONLY host emulation of the kernel body (see emu.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from letf_monte_carlo_b200 import abi
from oracle.oracle import HostOutputs

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
# LETF_EMU_SANITIZE=1 builds the emulation with AddressSanitizer + UBSan (run pytest under
# LD_PRELOAD=$(gcc -print-file-name=libasan.so)): out-of-bounds indexing of the kernel source's scratch / ring / history /
# tape accesses then aborts the test run - the CPU stand-in for compute-sanitizer
_SAN = os.environ.get("LETF_EMU_SANITIZE") == "1"
LIB = os.path.join(_BUILD, "libletf_emu_asan.so" if _SAN else "libletf_emu.so")
_lib = None


def build(force=False):
    os.makedirs(_BUILD, exist_ok=True)
    src = os.path.join(_HERE, "emu.cpp")
    csrc = os.path.join(_HERE, "..", "..", "letf_monte_carlo_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    deps.append(os.path.join(_HERE, "..", "..", "include", "letf_b200.h"))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"] if _SAN else []
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off"] + san + ["-o", LIB, src],
                       check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.letf_emu_run.argtypes = [C.POINTER(abi.Params), C.c_int, C.POINTER(abi.Tape), C.POINTER(abi.Outputs),
                                   C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int] + \
                                  [C.c_void_p] * 6
        L.letf_emu_run.restype = C.c_int
        L.letf_emu_run_lean.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Outputs), C.c_uint64, C.c_int64, C.c_int64,
                                        C.POINTER(C.c_int), C.c_int, C.c_int]
        L.letf_emu_run_lean.restype = C.c_int
        L.letf_emu_run_jpair.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Outputs), C.c_uint64, C.c_int64, C.c_int64,
                                         C.POINTER(C.c_int), C.c_int]
        L.letf_emu_run_jpair.restype = C.c_int
        L.letf_emu_run_typed.argtypes = [C.POINTER(abi.Params), C.c_int, C.POINTER(abi.Outputs), C.c_uint64, C.c_int64, C.c_int64]
        L.letf_emu_run_typed.restype = C.c_int
        L.letf_emu_market.argtypes = [C.POINTER(abi.Params), C.c_int, C.POINTER(abi.Tape), C.c_uint64, C.c_int64, C.c_int, C.c_void_p]
        L.letf_emu_market.restype = C.c_int
        L.letf_emu_run_split.argtypes = [C.POINTER(abi.Params), C.c_int, C.POINTER(abi.Outputs), C.c_uint64, C.c_int64, C.c_int64]
        L.letf_emu_run_split.restype = C.c_int
        L.letf_emu_run_series.argtypes = [C.POINTER(abi.Params), C.c_int, C.c_int, C.POINTER(abi.Series), C.POINTER(abi.Outputs),
                                          C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 7
        L.letf_emu_run_series.restype = C.c_int
        L.letf_emu_validate.argtypes = [C.POINTER(abi.Params)]
        L.letf_emu_validate.restype = C.c_char_p
        _lib = L
    return _lib


def run(params, tape, precision=abi.PREC_F64, rng_mode=abi.RNG_TAPE, seed=0, first=0, n=None, log=None):
    n_all = tape.n_tapes * (2 if params.antithetic_pairs else 1)
    n = n_all - first if n is None else n
    out = HostOutputs(params.n_strategies, n)
    ts = tape.as_struct()
    args = [-1, -1, 0] + [None] * 6
    if log is not None:
        trial, strat, cap = log
        bufs = {"count": np.zeros(1, np.int32), "day": np.zeros(cap, np.int32), "asset": np.zeros(cap, np.int32),
                "action": np.zeros(cap, np.int32), "shares": np.zeros(cap), "price": np.zeros(cap)}
        args = [trial, strat, cap] + [bufs[k].ctypes.data for k in ("count", "day", "asset", "action", "shares", "price")]
    rc = lib().letf_emu_run(C.byref(params), precision, C.byref(ts), C.byref(out.as_struct()), seed, first, n,
                            rng_mode, *args)
    if rc != 0:
        raise RuntimeError(f"emu rc={rc}: {lib().letf_emu_validate(C.byref(params))}")
    if log is not None:
        out.log = bufs
    return out


def run_series(params, asset_ret, *, tlt_ret=None, vix=None, regime=None, cash_ret=None, precision=abi.PREC_F64, typed=False,
               log=None, curve=False, split_list=None):
    """materialised-path mode on host series: asset_ret = {asset name: [T][n] array}; log = (path, strategy, cap) switches
    the journal log on, curve=True also returns that (path, strategy)'s daily equity (out.curve)"""
    cols = {k: np.ascontiguousarray(np.asarray(v, dtype=np.float64)) for k, v in asset_ret.items()}
    first = next(iter(cols.values()))
    T, n = first.shape
    s = abi.Series()
    s.n_paths, s.sim_days = n, T
    for name, a in cols.items():
        s.asset_ret[abi.ASSET_INDEX[name]] = a.ctypes.data
    keep = [np.ascontiguousarray(np.asarray(x, dtype=dt)) if x is not None else None
            for x, dt in ((tlt_ret, np.float64), (vix, np.float64), (regime, np.int32), (cash_ret, np.float64))]
    s.tlt_ret, s.vix, s.regime, s.cash_ret = (k.ctypes.data if k is not None else None for k in keep)
    out = HostOutputs(params.n_strategies, n)
    if split_list is not None:      # record pass + rule kernel body (letf_api.cu launch_series_split): 1 / 2 typed, 3 routed, 0 dynamic
        rc = lib().letf_emu_run_series_split(C.byref(params), split_list, C.byref(s), C.byref(out.as_struct()))
        if rc != 0:
            raise RuntimeError(f"emu series split rc={rc}")
        return out
    args = [-1, -1, 0] + [None] * 7
    if log is not None:
        path, strat, cap = log
        bufs = {"count": np.zeros(1, np.int32), "day": np.zeros(cap, np.int32), "asset": np.zeros(cap, np.int32),
                "action": np.zeros(cap, np.int32), "shares": np.zeros(cap), "price": np.zeros(cap)}
        cv = np.zeros(max(T, 1)) if curve else None
        args = [path, strat, cap] + [bufs[k].ctypes.data for k in ("count", "day", "asset", "action", "shares", "price")] + \
               [cv.ctypes.data if cv is not None else None]
    rc = lib().letf_emu_run_series(C.byref(params), precision, 1 if typed else 0, C.byref(s), C.byref(out.as_struct()), *args)
    if rc != 0:
        raise RuntimeError(f"emu series rc={rc}")
    if log is not None:
        out.log = bufs
        out.curve = cv[:T] if cv is not None else None
    return out


def run_lean(params, n, seed=0, first=0, two_per_thread=False, force_generic=False):
    """lean buy&hold build (NA = 3 compacted slots); returns (outputs, used_fast_variant).
    default: the straight-line fast path (letf_fast.cuh) with one trial per call when the shape admits it;
    two_per_thread: its launch form; force_generic: the lean build's general loop (Trial::run)"""
    out = HostOutputs(params.n_strategies, n)
    fast = C.c_int(0)
    rc = lib().letf_emu_run_lean(C.byref(params), C.byref(out.as_struct()), seed, first, n, C.byref(fast),
                                 1 if two_per_thread else 0, 1 if force_generic else 0)
    if rc != 0:
        raise RuntimeError(f"emu lean rc={rc}")
    return out, int(fast.value)


def run_jpair(params, n, seed=0, first=0, force_generic=False):
    """lean institutional buy&hold build (NA = 3 compacted slots, joint-t market); returns (outputs, pair-kernel shape code).
    default: the antithetic-pair body (letf_jfast.cuh) when the shape is admitted; force_generic: Trial::run"""
    out = HostOutputs(params.n_strategies, n)
    used = C.c_int(0)
    rc = lib().letf_emu_run_jpair(C.byref(params), C.byref(out.as_struct()), seed, first, n, C.byref(used),
                                  1 if force_generic else 0)
    if rc != 0:
        raise RuntimeError(f"emu jpair rc={rc}")
    return out, int(used.value)


def market_returns(params, tape, trial, precision=abi.PREC_F64, rng_mode=abi.RNG_TAPE, seed=0):
    """[sim_days][3] generated SPY / QQQ / TLT returns of one trial (the simulate_joint_returns_t seam)"""
    ret = np.zeros((max(int(params.sim_days), 1), 3))
    ts = tape.as_struct()
    rc = lib().letf_emu_market(C.byref(params), precision, C.byref(ts), seed, trial, rng_mode, ret.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"emu market rc={rc}")
    return ret[:int(params.sim_days)]


def run_split(params, slot_list, n, seed=0, first=0):
    """split strategy pipeline (pair producer -> per-day records -> typed rule kernel body), slot lists as run_typed"""
    out = HostOutputs(params.n_strategies, n)
    rc = lib().letf_emu_run_split(C.byref(params), slot_list, C.byref(out.as_struct()), seed, first, n)
    if rc != 0:
        raise RuntimeError(f"emu split rc={rc}")
    return out


def run_typed(params, slot_list, n, seed=0, first=0):
    """typed-slot build (csrc slot_list_type): 1 = config-4 sweep, 2 = the default 19-strategy sweep"""
    out = HostOutputs(params.n_strategies, n)
    rc = lib().letf_emu_run_typed(C.byref(params), slot_list, C.byref(out.as_struct()), seed, first, n)
    if rc != 0:
        raise RuntimeError(f"emu typed rc={rc}")
    return out
