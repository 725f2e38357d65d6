"""Trade journal seam (reference: letf/trade.py:6-94).

On the GPU path the journal is kept as running per-asset share positions inside
the fused kernel (entry count + dollar volume per trial) and, for the handful of
trials the host tax path samples (letf/reporting.py:457-467), replayed into the
reference's list-of-dicts `Trade_List` shape by `PathEngine.trade_list`.
This module keeps the importable names and the host-side record type.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from collections.abc import Sequence
from typing import Dict, List

ROTH_IDS = ["S1", "S2", "S3", "S4", "S5", "S6"]
TAXABLE_IDS = ["S%d" % i for i in range(7, 20)]


@dataclass
class Trade:
    day_index: int
    asset: str
    action: str
    shares: float
    price: float
    dollar_amount: float


class TradeJournal:
    """Share-based BUY/SELL log; semantics of letf/trade.py:20-73."""

    def __init__(self):
        self.trades: List[Trade] = []
        self.positions: Dict[str, float] = {}

    def log_allocation_change(self, day, asset, prev_allocation, new_allocation, portfolio_value, price):
        if price <= 0:
            return
        target = new_allocation * portfolio_value / price
        delta = target - self.positions.get(asset, 0.0)
        if abs(delta) < 0.001:
            return
        shares = delta if delta > 0 else abs(delta)
        self.trades.append(Trade(day, asset, "BUY" if delta > 0 else "SELL", shares, price, shares * price))
        self.positions[asset] = target

    def get_summary(self) -> dict:
        return {"count": len(self.trades), "volume": sum(t.dollar_amount for t in self.trades)}

    def get_full_trades(self) -> List[dict]:
        return [asdict(t) for t in self.trades]


class TradeList(Sequence):
    """A replayed `Trade_List` (trade.py:83-91: list of {day_index, asset, action, shares, price, dollar_amount}) kept as the
    replay's column arrays.  It reads like the reference's list of dicts - len(), indexing, slicing, iteration, == with a list -
    and builds a row's dict only when that row is asked for; the in-tree tax engine takes the columns as they are (thousands of
    trades per list at long horizons: the dict round trip was half of the tax table's time)."""
    __slots__ = ("day", "asset", "action", "shares", "price", "asset_names")

    def __init__(self, day, asset, action, shares, price, asset_names):
        self.day, self.asset, self.action, self.shares, self.price = day, asset, action, shares, price
        self.asset_names = tuple(asset_names)

    def __len__(self):
        return int(self.day.shape[0])

    def _row(self, i):
        s, p = float(self.shares[i]), float(self.price[i])
        return {"day_index": int(self.day[i]), "asset": self.asset_names[int(self.asset[i])],
                "action": "BUY" if int(self.action[i]) > 0 else "SELL", "shares": s, "price": p, "dollar_amount": s * p}

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._row(k) for k in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self._row(i)

    def __iter__(self):
        names = self.asset_names
        for d, a, t, s, p in zip(self.day.tolist(), self.asset.tolist(), self.action.tolist(), self.shares.tolist(), self.price.tolist()):
            yield {"day_index": d, "asset": names[a], "action": "BUY" if t > 0 else "SELL", "shares": s, "price": p, "dollar_amount": s * p}

    def __eq__(self, other):
        if isinstance(other, (list, tuple, TradeList)):
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        return NotImplemented

    __hash__ = None

    def __repr__(self):
        return f"TradeList({len(self)} trades)"
