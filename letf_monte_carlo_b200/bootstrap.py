"""Block pools of the legacy_hybrid bootstrap generator, as flat tables for the device.

Host-side mirror of `BlockBootstrapReturns.__init__` / `_create_regime_blocks`
(letf/simulation/bootstrap.py:27-160) and of `create_bootstrap_sampler` (:461-477, without the
pickle cache).  The reference keeps every 168-row block as its own copied array; blocks overlap
(stride 21), so here a pool entry is just (first history row, cumulative SPY return) over ONE
cleaned copy of the history, which is what the kernel gathers from.

Pool A ("economy"): every block; SPY is taken from these.  Pool B ("tech"): blocks that lie
entirely in the real-QQQ era (>= 1999-03-10); QQQ and TLT are taken from these.  Each pool is
split by the block's regime (median VIX < 25 -> 0).
"""
from __future__ import annotations

import numpy as np

BLOCK_STRIDE = 21                      # bootstrap.py:105
REAL_QQQ_CUTOFF = "1999-03-10"         # bootstrap.py:98


class BlockBootstrapPools:
    def __init__(self, df, block_size: int = 168):
        import pandas as pd
        self.block_size = bs = int(block_size)
        self.df = df                   # kept so that the same frame can be handed to the reference (tests / goldens)
        n = len(df)
        vix = np.asarray(df["VIX"].values, dtype=float)
        spy = np.asarray(df["SPY_Ret"].values, dtype=float)
        qqq = np.asarray(df["QQQ_Ret"].values, dtype=float) if "QQQ_Ret" in df.columns else spy * 1.25
        tlt = np.asarray(df["TLT_Ret"].values, dtype=float) if "TLT_Ret" in df.columns else spy * -0.25
        has_real_qqq = np.asarray(df.index >= pd.Timestamp(REAL_QQQ_CUTOFF))
        a = {0: [], 1: []}
        b = {0: [], 1: []}
        for start in range(0, n - bs + 1, BLOCK_STRIDE):
            end = start + bs
            bv = vix[start:end]
            with np.errstate(all="ignore"):
                med = np.nanmedian(bv) if np.isfinite(bv).any() else np.nan
            regime = 0 if med < 25 else 1
            bspy = spy[start:end]
            if np.isnan(bspy).sum() > bs // 4:
                continue
            ret = float(np.prod(1 + np.nan_to_num(bspy, nan=0)) - 1)
            a[regime].append((start, ret))
            if has_real_qqq[start:end].all():
                b[regime].append((start, ret))
        self.hist_spy = np.ascontiguousarray(np.nan_to_num(spy, nan=0))
        self.hist_qqq = np.ascontiguousarray(np.nan_to_num(qqq, nan=0))
        self.hist_tlt = np.ascontiguousarray(np.nan_to_num(tlt, nan=0))
        self.pool_a = {r: self._table(a[r]) for r in (0, 1)}
        self.pool_b = {r: self._table(b[r]) for r in (0, 1)}

    @staticmethod
    def _table(entries):
        start = np.ascontiguousarray(np.array([e[0] for e in entries], dtype=np.int32))
        ret = np.ascontiguousarray(np.array([e[1] for e in entries], dtype=np.float64))
        return start, ret

    @classmethod
    def from_reference(cls, sampler):
        """Rebuild the tables from a reference `BlockBootstrapReturns` (it keeps its frame and block size)."""
        return cls(sampler.df, sampler.block_size)

    def counts(self):
        return {"A": [len(self.pool_a[r][0]) for r in (0, 1)], "B": [len(self.pool_b[r][0]) for r in (0, 1)]}


def create_bootstrap_sampler(df, block_size: int | None = None, cfg=None) -> BlockBootstrapPools:
    """bootstrap.py:461-477: one sampler per run, block size cfg.BOOTSTRAP_BLOCK_SIZE."""
    if block_size is None:
        block_size = int(getattr(cfg, "BOOTSTRAP_BLOCK_SIZE", 168)) if cfg is not None else 168
    return BlockBootstrapPools(df, block_size)


def as_pools(sampler) -> BlockBootstrapPools:
    if isinstance(sampler, BlockBootstrapPools):
        return sampler
    if hasattr(sampler, "df") and hasattr(sampler, "block_size"):
        return BlockBootstrapPools.from_reference(sampler)
    raise TypeError("bootstrap_sampler must be a BlockBootstrapPools or the reference's BlockBootstrapReturns")


# correlation matrices of the two noise blends
SAMPLER_CORR = (  # bootstrap.py:404-413 (factored as given)
    np.array([[1.000, 0.835, -0.207], [0.835, 1.000, -0.150], [-0.207, -0.150, 1.000]]),
    np.array([[1.000, 0.950, -0.447], [0.950, 1.000, -0.400], [-0.447, -0.400, 1.000]]),
)
ENGINE_CORR = (   # engine.py:60-61 (projected with nearest_psd_matrix first)
    np.array([[1.0, 0.88, -0.18], [0.88, 1.0, -0.12], [-0.18, -0.12, 1.0]]),
    np.array([[1.0, 0.94, -0.42], [0.94, 1.0, -0.30], [-0.42, -0.30, 1.0]]),
)
