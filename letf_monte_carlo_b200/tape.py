"""Shock tapes: pre-drawn unit-scale variates in the engine's day-major layout.

A tape is the parity-mode input of the fused kernel (LETF_RNG_TAPE) and the only
input of the CPU oracle.  `make_tape` draws one from a seeded numpy Generator;
the draw ledger it covers is SURVEY.md Appendix B (reference call sites are
listed next to each field in include/letf_b200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

F64_FIELDS = ("x", "chi2", "vix_n", "irx_n", "liq_n", "cred_n", "jump_u", "jump_t", "sign", "te_t", "boot_u", "boot_t",
              "boot_syn")


class HostTape:
    """numpy-backed LetfTape."""

    def __init__(self, n_tapes: int, gen_days: int, sim_days: int, max_spells: int, max_blocks: int = 0,
                 syn_block_size: int = 0):
        n, G, T, K = int(n_tapes), int(gen_days), int(sim_days), int(max_spells)
        self.n_tapes, self.gen_days, self.sim_days, self.max_spells = n, G, T, K
        self.max_blocks = B = int(max_blocks)         # MARKET_BOOTSTRAP only
        self.boot_len = np.ones((B, n), np.int32)
        self.boot_u = np.zeros((B, 6, n))
        self.boot_t = np.zeros((G if B else 0, 3, n))
        self.boot_block_size = int(syn_block_size) if B else 0   # > 0: some pool is empty -> synthetic-block variates
        self.boot_syn = np.zeros((B if self.boot_block_size else 0, 2, self.boot_block_size, n))
        self.start_offset = np.zeros(n, np.int32)
        self.start_regime = np.zeros(n, np.int32)
        self.spells = np.full((K, n), max(G, 1), np.int32)
        self.x = np.zeros((G, 3, n))
        self.chi2 = np.ones((G, n))
        self.vix_n = np.zeros((G, n))
        self.irx_n = np.zeros((G, n))
        self.liq_n = np.zeros((T, n))
        self.cred_n = np.zeros((T, n))
        self.jump_u = np.ones((T, n))
        self.jump_t = np.zeros((T, n))
        self.sign = np.ones((T, n))
        self.te_t = np.zeros((abi.N_ASSETS, T, n))

    def arrays(self):
        return {k: getattr(self, k) for k in abi.TAPE_FIELDS}

    def as_struct(self) -> abi.Tape:
        t = abi.Tape()
        t.n_tapes, t.gen_days, t.sim_days, t.max_spells = self.n_tapes, self.gen_days, self.sim_days, self.max_spells
        t.max_blocks = self.max_blocks
        t.boot_block_size = self.boot_block_size
        for k, a in self.arrays().items():
            assert a.flags["C_CONTIGUOUS"]
            setattr(t, k, a.ctypes.data if a.size else None)
        return t

    def regime_full(self, params: abi.Params) -> np.ndarray:
        """[G][n] regime path implied by the spells (engine.py:255-302 with 2 regimes)."""
        G, n = self.gen_days, self.n_tapes
        reg = np.zeros((G, n), np.int32)
        if params.market_mode == abi.MARKET_NORMAL:
            return reg
        flip = [params.transition[0][1] > 0, params.transition[1][0] > 0]
        for c in range(n):
            cur, t, k = int(self.start_regime[c]), 0, 0
            while t < G:
                sp = int(self.spells[k, c]) if k < self.max_spells else G
                sp = max(sp, 1)
                k += 1
                reg[t:t + sp, c] = cur
                t += sp
                if flip[cur]:
                    cur = 1 - cur
        return reg


def boot_tape_shape(params: abi.Params):
    """(max_blocks, synthetic block size) of a tape for these parameters: (0, 0) outside MARKET_BOOTSTRAP; the
    synthetic-block array only exists when some pool is empty (bootstrap.py:214-222)."""
    if params.market_mode != abi.MARKET_BOOTSTRAP:
        return 0, 0
    b = params.boot
    G = int(params.sim_days + params.buffer_days)
    empty = any(pl[r].n_blocks == 0 for pl in (b.pool_a, b.pool_b) for r in range(2))
    return G // max(int(b.block_min), 1) + 2, (int(b.block_size) if empty else 0)


def make_tape(params: abi.Params, n_tapes: int, seed: int) -> HostTape:
    """Draw a tape with the distributions the reference would use for these parameters."""
    rng = np.random.default_rng(seed)
    n, T = int(n_tapes), int(params.sim_days)
    G = T + int(params.buffer_days)
    normal_mode = params.market_mode == abi.MARKET_NORMAL
    start_regime = (rng.random(n) < params.start_regime_prob1).astype(np.int32) if params.start_regime_prob1 > 0 \
        else np.zeros(n, np.int32)
    # ---- dwell-time draws until every column covers G days
    spells, cover, k = [], np.zeros(n, np.int64), 0
    if not normal_mode:
        while cover.min() < G:
            # regime of spell k per column (both regimes flip in every test model; otherwise the path stays put)
            both = params.transition[0][1] > 0 and params.transition[1][0] > 0
            rcol = (start_regime + k) % 2 if both else start_regime
            s = np.zeros(n, np.int64)
            for r in (0, 1):
                m = rcol == r
                if not m.any():
                    continue
                if params.n_duration[r] > 0:
                    pool = np.ctypeslib.as_array(params.duration_samples[r], shape=(params.n_duration[r],))
                    s[m] = rng.choice(pool, size=int(m.sum()))
                else:
                    p_stay = float(np.clip(params.transition[r][r], 0.80, 0.995))
                    s[m] = np.maximum(1, rng.geometric(1 - p_stay, size=int(m.sum())))
            spells.append(s.astype(np.int32))
            cover += np.maximum(s, 1)
            k += 1
    boot = params.market_mode == abi.MARKET_BOOTSTRAP
    tp = HostTape(n, G, T, max(len(spells), 1), *boot_tape_shape(params))
    for i, s in enumerate(spells):
        tp.spells[i] = s
    tp.start_regime[:] = start_regime
    if params.random_start and params.buffer_days > 0:
        b = int(params.buffer_days)
        tp.start_offset[:] = rng.integers(min(50, b // 4), b, size=n)
    reg = tp.regime_full(params)
    if boot:       # bootstrap.py:302-308 block lengths, :212-262 selection uniforms, :400 blend noise
        tp.boot_len[:] = rng.geometric(1.0 / params.boot.block_mean, size=tp.boot_len.shape)
        tp.boot_u[:] = rng.random(tp.boot_u.shape)
        tp.boot_t[:] = rng.standard_t(params.param_t_df, size=tp.boot_t.shape)
        if tp.boot_syn.size:
            tp.boot_syn[:] = rng.standard_t(params.param_t_df, size=tp.boot_syn.shape)
    if params.market_mode in (abi.MARKET_PARAM_T, abi.MARKET_BOOTSTRAP):
        tp.x[:] = rng.standard_t(params.param_t_df, size=tp.x.shape)
    else:
        tp.x[:] = rng.standard_normal(tp.x.shape)
    if params.market_mode == abi.MARKET_JOINT_T:
        nu = np.where(reg == 0, params.joint[0].nu, params.joint[1].nu)
        tp.chi2[:] = rng.chisquare(nu)
    if params.has_vix_dynamics:
        tp.vix_n[:] = rng.standard_normal((G, n))
        tp.irx_n[:] = rng.standard_normal((G, n))
    if params.has_stress:
        tp.liq_n[:] = rng.standard_normal((T, n))
        tp.cred_n[:] = rng.standard_normal((T, n))
        tp.jump_u[:] = rng.random((T, n))
        tp.jump_t[:] = rng.standard_t(5, size=(T, n))
        tp.sign[:] = rng.choice([-1.0, 1.0], size=(T, n))
    for a in range(abi.N_ASSETS):
        if (params.asset_mask >> a) & 1:
            tp.te_t[a] = rng.standard_t(params.assets[a].te_df, size=(T, n))
    return tp
