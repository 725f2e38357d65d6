"""letf_monte_carlo_b200 - B200-native LETF Monte Carlo path engine.

Drop-in for the hot path of djbm10/LETF-Monte-Carlo
(`letf.mc_runner.parallel_monte_carlo_fixed` -> `simulate_single_path_fixed` ->
`run_strategy_fixed`); everything else of the reference is out of scope.
"""
from . import abi, config  # noqa: F401
from .calibration import simulate_joint_returns_t  # noqa: F401
from .mc_runner import parallel_monte_carlo_fixed  # noqa: F401
from .simulation import simulate_single_path_fixed  # noqa: F401
from .strategy import run_strategy_fixed  # noqa: F401
from .trade import TAXABLE_IDS, ROTH_IDS, TradeJournal  # noqa: F401

__all__ = ["abi", "config", "parallel_monte_carlo_fixed", "simulate_single_path_fixed", "run_strategy_fixed", "simulate_joint_returns_t",
           "TAXABLE_IDS", "ROTH_IDS", "TradeJournal"]
