"""ctypes binding of include/finrl_b200.h — the only door between Python and the CUDA engine.

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception
is raised.  Nothing in this package computes an env step on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# FINRL_B200_LIB overrides the library path (used for A/B builds of the same ABI while tuning)
LIB_PATH = os.environ.get("FINRL_B200_LIB") or os.path.join(PKG, "libfinrl_b200.so")

FLAG_DONE = 1
FLAG_LIQUIDATE = 2
FLAG_SHORTAGE = 4

OBS_NONE, OBS_LAST, OBS_ALL = 0, 1, 2
N_STATS = 8
STAT_NAMES = ("reward_sum", "reward_sqsum", "done_count", "episode_asset_sum", "asset_sum", "liq_count", "env_steps",
              "trades_sum")

ABI_VERSION = 4
STATS_BLOCK_DOUBLES = 48  # frl_stats_block: 384 bytes = sum[2][8], total[8], peer_total[8], n_peers, reserved


class EngineError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


class TradingParams(C.Structure):
    """frl_trading_params (include/finrl_b200.h)."""

    _fields_ = [
        ("n_envs", C.c_int32),
        ("stock_dim", C.c_int32),
        ("n_tech", C.c_int32),
        ("n_days", C.c_int32),
        ("obs_dim", C.c_int32),
        ("env_stride", C.c_int32),
        ("hmax", C.c_double),
        ("initial_amount", C.c_double),
        ("buy_cost_pct", C.c_double),
        ("sell_cost_pct", C.c_double),
        ("reward_scaling", C.c_double),
        ("use_turbulence", C.c_int32),
        ("close_pitch", C.c_int32),
        ("turbulence_threshold", C.c_double),
        ("close", C.c_void_p),
        ("disable_mask", C.c_void_p),
        ("risk", C.c_void_p),
        ("obs_tmpl", C.c_void_p),
        ("init_hold", C.c_void_p),
        ("cash", C.c_void_p),
        ("hold", C.c_void_p),
        ("day", C.c_void_p),
        ("sday", C.c_void_p),
        ("cost", C.c_void_p),
        ("trades", C.c_void_p),
        ("reward", C.c_void_p),
        ("episode", C.c_void_p),
        ("asset_out", C.c_void_p),
        ("obs_tmpl4", C.c_void_p),
    ]


class NpParams(C.Structure):
    """frl_np_params (include/finrl_b200.h)."""

    _fields_ = [
        ("n_envs", C.c_int32),
        ("stock_dim", C.c_int32),
        ("tech_dim", C.c_int32),
        ("n_days", C.c_int32),
        ("obs_dim", C.c_int32),
        ("env_stride", C.c_int32),
        ("gamma", C.c_double),
        ("max_stock", C.c_double),
        ("min_stock_rate", C.c_double),
        ("buy_cost_pct", C.c_double),
        ("sell_cost_pct", C.c_double),
        ("reward_scaling", C.c_double),
        ("initial_capital", C.c_double),
        ("obs_amount_floor", C.c_double),
        ("price", C.c_void_p),
        ("turb_bool", C.c_void_p),
        ("obs_tmpl", C.c_void_p),
        ("init_stocks", C.c_void_p),
        ("amount", C.c_void_p),
        ("kinds", C.c_void_p),
        ("stocks", C.c_void_p),
        ("cool", C.c_void_p),
        ("day", C.c_void_p),
        ("total", C.c_void_p),
        ("gamma_reward", C.c_void_p),
        ("init_total", C.c_void_p),
        ("episode_return", C.c_void_p),
        ("price_pitch", C.c_int32),
        ("train_reset", C.c_int32),
        ("reset_seed", C.c_uint64),
    ]


class PortfolioParams(C.Structure):
    """frl_portfolio_params (include/finrl_b200.h)."""

    _fields_ = [
        ("n_envs", C.c_int32),
        ("stock_dim", C.c_int32),
        ("n_tech", C.c_int32),
        ("n_days", C.c_int32),
        ("obs_dim", C.c_int32),
        ("_pad0", C.c_int32),
        ("initial_amount", C.c_double),
        ("ret", C.c_void_p),
        ("obs_table", C.c_void_p),
        ("pv", C.c_void_p),
        ("day", C.c_void_p),
        ("reward", C.c_void_p),
        ("ret_out", C.c_void_p),
        ("weights_out", C.c_void_p),
        ("ret_pitch", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class StopLossParams(C.Structure):
    """frl_stoploss_params (include/finrl_b200.h)."""

    _fields_ = [
        ("n_envs", C.c_int32),
        ("stock_dim", C.c_int32),
        ("n_cols", C.c_int32),
        ("n_days", C.c_int32),
        ("obs_dim", C.c_int32),
        ("discrete_actions", C.c_int32),
        ("shares_increment", C.c_int32),
        ("use_turbulence", C.c_int32),
        ("patient", C.c_int32),
        ("env_stride", C.c_int32),
        ("buy_cost_pct", C.c_double),
        ("sell_cost_pct", C.c_double),
        ("hmax", C.c_double),
        ("turbulence_threshold", C.c_double),
        ("initial_amount", C.c_double),
        ("cash_penalty_proportion", C.c_double),
        ("stoploss_penalty", C.c_double),
        ("min_profit_penalty", C.c_double),
        ("close", C.c_void_p),
        ("turb", C.c_void_p),
        ("obs_tmpl", C.c_void_p),
        ("cash", C.c_void_p),
        ("assets", C.c_void_p),
        ("date_index", C.c_void_p),
        ("start", C.c_void_p),
        ("fresh", C.c_void_p),
        ("last_cash", C.c_void_p),
        ("last_total", C.c_void_p),
        ("sum_trades", C.c_void_p),
        ("hmax_vec", C.c_void_p),
        ("hmax_vec_f32", C.c_int32),
        ("random_start", C.c_int32),
        ("reset_seed", C.c_uint64),
    ]


class CashPenaltyParams(C.Structure):
    """frl_cashpenalty_params (include/finrl_b200.h)."""

    _fields_ = [
        ("n_envs", C.c_int32),
        ("stock_dim", C.c_int32),
        ("n_cols", C.c_int32),
        ("n_days", C.c_int32),
        ("obs_dim", C.c_int32),
        ("discrete_actions", C.c_int32),
        ("shares_increment", C.c_int32),
        ("use_turbulence", C.c_int32),
        ("patient", C.c_int32),
        ("env_stride", C.c_int32),
        ("buy_cost_pct", C.c_double),
        ("sell_cost_pct", C.c_double),
        ("hmax", C.c_double),
        ("turbulence_threshold", C.c_double),
        ("initial_amount", C.c_double),
        ("cash_penalty_proportion", C.c_double),
        ("close", C.c_void_p),
        ("turb", C.c_void_p),
        ("obs_tmpl", C.c_void_p),
        ("cash", C.c_void_p),
        ("hold", C.c_void_p),
        ("hold_alt", C.c_void_p),
        ("date_index", C.c_void_p),
        ("start", C.c_void_p),
        ("fresh", C.c_void_p),
        ("last_cash", C.c_void_p),
        ("last_total", C.c_void_p),
        ("sum_trades", C.c_void_p),
        ("hmax_vec", C.c_void_p),
        ("hmax_vec_f32", C.c_int32),
        ("random_start", C.c_int32),
        ("reset_seed", C.c_uint64),
        ("close_rc", C.c_void_p),
    ]


class CryptoParams(C.Structure):
    """frl_crypto_params (include/finrl_b200.h)."""

    _fields_ = [
        ("n_envs", C.c_int32),
        ("stock_dim", C.c_int32),
        ("tech_dim", C.c_int32),
        ("n_days", C.c_int32),
        ("lookback", C.c_int32),
        ("obs_dim", C.c_int32),
        ("env_stride", C.c_int32),
        ("_pad0", C.c_int32),
        ("initial_capital", C.c_double),
        ("buy_cost_pct", C.c_double),
        ("sell_cost_pct", C.c_double),
        ("gamma", C.c_double),
        ("price", C.c_void_p),
        ("act_norm", C.c_void_p),
        ("obs_tmpl", C.c_void_p),
        ("cash", C.c_void_p),
        ("stocks", C.c_void_p),
        ("time", C.c_void_p),
        ("total", C.c_void_p),
        ("gamma_return", C.c_void_p),
        ("episode_return", C.c_void_p),
    ]


KIND_PY, KIND_F32, KIND_F64 = 0, 1, 2
NP_REWARD_KIND_SHIFT = 4

# name -> (restype, argtypes); every symbol include/finrl_b200.h declares
SIGNATURES = {
    "frl_abi_version": (C.c_int32, []),
    "frl_last_error": (C.c_char_p, []),
    "frl_set_option": (C.c_int32, [C.c_char_p, C.c_int64]),
    "frl_trading_init": (C.c_int32, [C.POINTER(TradingParams), C.c_int32, C.c_void_p]),
    "frl_trading_reset": (C.c_int32, [C.POINTER(TradingParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_trading_observe": (C.c_int32, [C.POINTER(TradingParams), C.c_void_p, C.c_void_p]),
    "frl_trading_observe_factored": (C.c_int32, [C.POINTER(TradingParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_expand_obs_host": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_int32]),
    "frl_expand_obs_host_chunks": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                               C.c_int32]),
    "frl_trading_rollout": (
        C.c_int32,
        [C.POINTER(TradingParams), C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "frl_trading_step": (
        C.c_int32,
        [C.POINTER(TradingParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
         C.c_void_p],
    ),
    "frl_np_reset": (C.c_int32, [C.POINTER(NpParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_np_observe": (C.c_int32, [C.POINTER(NpParams), C.c_void_p, C.c_void_p]),
    "frl_np_rollout": (
        C.c_int32,
        [C.POINTER(NpParams), C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "frl_np_step": (
        C.c_int32,
        [C.POINTER(NpParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
         C.c_void_p],
    ),
    "frl_portfolio_reset": (C.c_int32, [C.POINTER(PortfolioParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_portfolio_observe": (C.c_int32, [C.POINTER(PortfolioParams), C.c_void_p, C.c_void_p]),
    "frl_portfolio_rollout": (
        C.c_int32,
        [C.POINTER(PortfolioParams), C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "frl_portfolio_step": (
        C.c_int32,
        [C.POINTER(PortfolioParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
         C.c_void_p],
    ),
    "frl_cashpenalty_reset": (C.c_int32, [C.POINTER(CashPenaltyParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_cashpenalty_observe": (C.c_int32, [C.POINTER(CashPenaltyParams), C.c_void_p, C.c_void_p]),
    "frl_cashpenalty_rollout": (
        C.c_int32,
        [C.POINTER(CashPenaltyParams), C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "frl_cashpenalty_step": (
        C.c_int32,
        [C.POINTER(CashPenaltyParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
         C.c_void_p],
    ),
    "frl_stoploss_reset": (C.c_int32, [C.POINTER(StopLossParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_stoploss_observe": (C.c_int32, [C.POINTER(StopLossParams), C.c_void_p, C.c_void_p]),
    "frl_stoploss_rollout": (
        C.c_int32,
        [C.POINTER(StopLossParams), C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "frl_stoploss_step": (
        C.c_int32,
        [C.POINTER(StopLossParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
         C.c_void_p],
    ),
    "frl_crypto_reset": (C.c_int32, [C.POINTER(CryptoParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_crypto_observe": (C.c_int32, [C.POINTER(CryptoParams), C.c_void_p, C.c_void_p]),
    "frl_crypto_rollout": (
        C.c_int32,
        [C.POINTER(CryptoParams), C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "frl_crypto_step": (
        C.c_int32,
        [C.POINTER(CryptoParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
         C.c_void_p],
    ),
    "frl_exchange_alloc": (C.c_int32, [C.POINTER(C.c_void_p)]),
    "frl_exchange_free": (C.c_int32, [C.c_void_p]),
    "frl_exchange_export": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "frl_exchange_open": (C.c_int32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "frl_exchange_close": (C.c_int32, [C.c_void_p]),
    "frl_exchange_bind": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "frl_exchange_flush": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "frl_turbulence": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "frl_rolling_cov": (
        C.c_int32,
        [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
}

_lib = None


def lib():
    """Load the shared library once; fail loudly when it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} not found: build it with `python -m finrl_b200.build` "
                "(nvcc, sm_100a).  finrl_b200 has no CPU fallback."
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if l.frl_abi_version() != ABI_VERSION:
            raise EngineError(f"ABI version mismatch: library {l.frl_abi_version()}, binding {ABI_VERSION}")
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().frl_last_error().decode("utf-8", "replace")
        raise EngineError(f"{what} failed (status {rc}): {msg}")


def set_option(name: str, value: int):
    """frl_set_option: e.g. set_option("trading_small_max", 0) forces the thread-per-env trading kernel."""
    check(lib().frl_set_option(name.encode(), int(value)), f"frl_set_option({name})")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream(device):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def new_stats_block(torch, device):
    """A zeroed frl_stats_block (48 doubles, its own 512-byte-aligned torch allocation): ``block[:8]`` and
    ``block[8:16]`` are the two accumulators the kernels add into, ``block[16:24]`` the exchange totals."""
    return torch.zeros(STATS_BLOCK_DOUBLES, dtype=torch.float64, device=device)


class _DevicePointerView:
    """``__cuda_array_interface__`` over a raw device pointer, so torch can view memory this library
    allocated with cudaMalloc (the IPC-exportable statistics blocks)."""

    def __init__(self, pointer: int, n_doubles: int):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (int(pointer), False),
                                         "version": 2, "strides": None}
