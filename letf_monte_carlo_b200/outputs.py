"""Per-trial result buffers in the engine's structure-of-arrays layout ([strategy][trial]).

`HostOutputs` is numpy-backed (what `letf_run_host` fills), `DeviceOutputs` is backed by
torch CUDA tensors (what `letf_run` fills; torch is used for device memory only).
Field meaning: include/letf_b200.h `LetfOutputs` (reference: engine.py:578-603).
"""
from __future__ import annotations

import numpy as np

from . import abi

FIELDS = ("final_wealth", "max_dd", "num_trades", "journal_count", "journal_volume",
          "start_offset", "flags", "regime_days", "jump_days")
_OPTIONAL_JOURNAL = ("journal_count", "journal_volume")


def _host_array(shape, dtype, pinned):
    if pinned:
        import torch
        tdt = {np.float64: torch.float64, np.int32: torch.int32, np.uint32: torch.int32}[dtype]
        a = torch.zeros(shape, dtype=tdt, pin_memory=True).numpy()      # page-locked: device->host at full PCIe rate
        return a.view(np.uint32) if dtype is np.uint32 else a
    return np.zeros(shape, dtype)


class HostOutputs:
    def __init__(self, n_strategies: int, n: int, journal: bool = True, pinned: bool = False):
        S = max(int(n_strategies), 1)
        self.n = int(n)
        self.n_strategies = int(n_strategies)
        z = lambda shape, dt: _host_array(shape, dt, pinned)
        self.final_wealth = z((S, n), np.float64)
        self.max_dd = z((S, n), np.float64)
        self.num_trades = z((S, n), np.int32)
        self.journal_count = z((S, n), np.int32) if journal else None
        self.journal_volume = z((S, n), np.float64) if journal else None
        self.start_offset = z(n, np.int32)
        self.flags = z(n, np.uint32)
        self.regime_days = z((2, n), np.int32)
        self.jump_days = z(n, np.int32)

    def as_struct(self) -> abi.Outputs:
        o = abi.Outputs()
        o.n_trials = self.n
        for k in FIELDS:
            a = getattr(self, k)
            setattr(o, k, a.ctypes.data if a is not None else None)
        return o

    def as_dict(self):
        return {k: getattr(self, k) for k in FIELDS if getattr(self, k) is not None}


class DeviceOutputs:
    def __init__(self, n_strategies: int, n: int, device, journal: bool = False):
        import torch
        S = max(int(n_strategies), 1)
        self.n = int(n)
        self.n_strategies = int(n_strategies)
        kw = {"device": device}
        self.final_wealth = torch.zeros((S, n), dtype=torch.float64, **kw)
        self.max_dd = torch.zeros((S, n), dtype=torch.float64, **kw)
        self.num_trades = torch.zeros((S, n), dtype=torch.int32, **kw)
        self.journal_count = torch.zeros((S, n), dtype=torch.int32, **kw) if journal else None
        self.journal_volume = torch.zeros((S, n), dtype=torch.float64, **kw) if journal else None
        self.start_offset = torch.zeros(n, dtype=torch.int32, **kw)
        self.flags = torch.zeros(n, dtype=torch.int32, **kw)        # bit pattern of the uint32 flags
        self.regime_days = torch.zeros((2, n), dtype=torch.int32, **kw)
        self.jump_days = torch.zeros(n, dtype=torch.int32, **kw)

    def as_struct(self) -> abi.Outputs:
        o = abi.Outputs()
        o.n_trials = self.n
        for k in FIELDS:
            t = getattr(self, k)
            setattr(o, k, t.data_ptr() if t is not None else None)
        return o

    def to_host(self) -> HostOutputs:
        h = HostOutputs(self.n_strategies, self.n, journal=self.journal_count is not None)
        for k in FIELDS:
            t = getattr(self, k)
            if t is None:
                continue
            a = t.cpu().numpy()
            getattr(h, k)[...] = a.view(np.uint32) if k == "flags" else a
        return h
