"""finrl_b200 — B200-native batched trading-environment engine (drop-in for FinRL's env step path).

Host side: Python + PyTorch tensors (device memory, streams, torch.distributed).
Compute: hand-written sm_100a CUDA behind the C-ABI of include/finrl_b200.h
(finrl_b200/libfinrl_b200.so).  There is no CPU fallback: importing the env classes works
anywhere, constructing one without the built library or a CUDA device raises.
"""
from ._cabi import EngineError, FLAG_DONE, FLAG_LIQUIDATE, FLAG_SHORTAGE  # noqa: F401
from .tables import TradingTables, frame_to_arrays  # noqa: F401
from .trading import BatchedStockTradingEnv  # noqa: F401
from .nptrading import BatchedNpStockTradingEnv, NpTables  # noqa: F401
from .portfolio import BatchedStockPortfolioEnv, PortfolioTables  # noqa: F401
from .cashpenalty import BatchedStockTradingEnvCashpenalty, CashPenaltyTables  # noqa: F401
from .stoploss import BatchedStockTradingEnvStopLoss  # noqa: F401

__version__ = "0.1.0"
