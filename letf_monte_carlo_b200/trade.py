"""Trade journal seam (reference: letf/trade.py:6-94).

On the GPU path the journal is kept as running per-asset share positions inside
the fused kernel (entry count + dollar volume per trial) and, for the handful of
trials the host tax path samples (letf/reporting.py:457-467), replayed into the
reference's list-of-dicts `Trade_List` shape by `PathEngine.trade_list`.
This module keeps the importable names and the host-side record type.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
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
