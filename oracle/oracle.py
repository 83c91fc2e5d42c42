"""TEST INFRASTRUCTURE — ctypes front-end of the C oracle (oracle/oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package ``finrl_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_DONE = 1
FLAG_LIQUIDATE = 2
FLAG_SHORTAGE = 4


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ora_floor_divide_f64.restype = C.c_double
        _LIB.ora_floor_divide_f64.argtypes = [C.c_double, C.c_double]
        _LIB.ora_floor_divide_f32.restype = C.c_float
        _LIB.ora_floor_divide_f32.argtypes = [C.c_float, C.c_float]
        _LIB.ora_pairwise_sum_f32.restype = C.c_float
        _LIB.ora_pairwise_sum_f64.restype = C.c_double
    return _LIB


def set_threads(n: int) -> int:
    """Force the OpenMP thread count of the env-parallel loops (torchrun exports OMP_NUM_THREADS=1)."""
    return int(lib().ora_set_threads(C.c_int(int(n))))


def _p(a, ty=C.c_void_p):
    return a.ctypes.data_as(ty) if a is not None else None


def floor_divide_f64(a, b):
    return lib().ora_floor_divide_f64(float(a), float(b))


def floor_divide_f32(a, b):
    return lib().ora_floor_divide_f32(float(a), float(b))


def argsort_i64(keys):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    out = np.empty(keys.shape[0], dtype=np.int32)
    lib().ora_argsort_i64(_p(keys), C.c_int(keys.shape[0]), _p(out))
    return out


def pairwise_sum_f32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return np.float32(lib().ora_pairwise_sum_f32(_p(x), C.c_int(x.shape[0])))


def pairwise_sum_f64(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return np.float64(lib().ora_pairwise_sum_f64(_p(x), C.c_int(x.shape[0])))


# --------------------------------------------------------------------------------------------
# A1
# --------------------------------------------------------------------------------------------
class _TradingCfg(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32),
        ("stock_dim", C.c_int32),
        ("n_tech", C.c_int32),
        ("n_days", C.c_int32),
        ("hmax", C.c_double),
        ("initial_amount", C.c_double),
        ("buy_cost_pct", C.c_double),
        ("sell_cost_pct", C.c_double),
        ("reward_scaling", C.c_double),
        ("use_turbulence", C.c_int32),
        ("turbulence_threshold", C.c_double),
        ("close", C.c_void_p),
        ("tech", C.c_void_p),
        ("risk", C.c_void_p),
        ("init_hold", C.c_void_p),
    ]


class _TradingState(C.Structure):
    _fields_ = [
        ("cash", C.c_void_p),
        ("hold", C.c_void_p),
        ("day", C.c_void_p),
        ("sday", C.c_void_p),
        ("cost", C.c_void_p),
        ("trades", C.c_void_p),
        ("reward", C.c_void_p),
        ("episode", C.c_void_p),
    ]


class TradingOracle:
    """N independent copies of the reference ``StockTradingEnv``
    (finrl/meta/env_stock_trading/env_stocktrading.py), stepped on the CPU."""

    def __init__(
        self,
        close,
        tech,
        risk,
        n_envs,
        hmax=100,
        initial_amount=1_000_000,
        buy_cost_pct=0.001,
        sell_cost_pct=0.001,
        reward_scaling=1e-4,
        turbulence_threshold=None,
        num_stock_shares=None,
        day=0,
    ):
        self.close = np.ascontiguousarray(close, dtype=np.float64)
        T, D = self.close.shape
        self.tech = np.ascontiguousarray(tech, dtype=np.float64).reshape(-1, T, D)
        K = self.tech.shape[0]
        self.risk = np.ascontiguousarray(risk if risk is not None else np.zeros(T), dtype=np.float64)
        self.N, self.D, self.K, self.T = int(n_envs), D, K, T
        self.O = 1 + 2 * D + K * D
        self.init_hold = np.ascontiguousarray(
            num_stock_shares if num_stock_shares is not None else np.zeros(D), dtype=np.int32
        )
        N = self.N
        self.cash = np.zeros(N)
        self.hold = np.zeros((N, D), dtype=np.int32)
        self.day = np.zeros(N, dtype=np.int32)
        self.sday = np.zeros(N, dtype=np.int32)
        self.cost = np.zeros(N)
        self.trades = np.zeros(N, dtype=np.int32)
        self.reward = np.zeros(N)
        self.episode = np.zeros(N, dtype=np.int32)
        self._cfg = _TradingCfg(
            N, D, K, T, float(hmax), float(initial_amount), float(buy_cost_pct), float(sell_cost_pct),
            float(reward_scaling), int(turbulence_threshold is not None),
            float(turbulence_threshold if turbulence_threshold is not None else 0.0),
            _p(self.close), _p(self.tech), _p(self.risk), _p(self.init_hold),
        )
        self._st = _TradingState(
            _p(self.cash), _p(self.hold), _p(self.day), _p(self.sday), _p(self.cost), _p(self.trades),
            _p(self.reward), _p(self.episode),
        )
        lib().ora_trading_init(C.byref(self._cfg), C.byref(self._st), C.c_int32(day))

    def obs(self):
        out = np.empty((self.N, self.O), dtype=np.float32)
        lib().ora_trading_obs(C.byref(self._cfg), C.byref(self._st), _p(out))
        return out

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().ora_trading_reset(C.byref(self._cfg), C.byref(self._st), _p(m))
        return self.obs()

    def step(self, actions, auto_reset=False, want_obs=True):
        a = np.ascontiguousarray(actions)
        assert a.shape == (self.N, self.D) and a.dtype in (np.float32, np.float64)
        reward = np.empty(self.N)
        flags = np.empty(self.N, dtype=np.uint8)
        obs = np.empty((self.N, self.O), dtype=np.float32) if want_obs else None
        lib().ora_trading_step(
            C.byref(self._cfg), C.byref(self._st), _p(a), C.c_int(a.dtype == np.float64), _p(reward), _p(flags),
            _p(obs), C.c_int(int(auto_reset)),
        )
        return obs, reward, flags


# --------------------------------------------------------------------------------------------
# A2
# --------------------------------------------------------------------------------------------
KIND_PY, KIND_F32, KIND_F64 = 0, 1, 2


def np_tables(price_array, tech_array, turbulence_array, turbulence_thresh=99):
    """The table transforms of the numpy env's constructor (env_stocktrading_np.py:27-35, 164-169)."""
    price = np.ascontiguousarray(price_array.astype(np.float32))
    tech = np.ascontiguousarray(tech_array.astype(np.float32) * 2**-7)
    turb_bool = (turbulence_array > turbulence_thresh).astype(np.float32)

    def sigmoid(x):
        return 1 / (1 + np.exp(-x * np.e)) - 0.5

    turb = (sigmoid(turbulence_array / turbulence_thresh) * turbulence_thresh * 2**-5).astype(np.float32)
    return price, tech, turb_bool, turb


class _NpCfg(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("stock_dim", C.c_int32), ("tech_dim", C.c_int32), ("n_days", C.c_int32),
        ("gamma", C.c_double), ("max_stock", C.c_double), ("min_stock_rate", C.c_double),
        ("buy_cost_pct", C.c_double), ("sell_cost_pct", C.c_double), ("reward_scaling", C.c_double),
        ("initial_capital", C.c_double), ("obs_amount_floor", C.c_double),
        ("price", C.c_void_p), ("tech", C.c_void_p), ("turb_bool", C.c_void_p), ("turb_ary", C.c_void_p),
        ("init_stocks", C.c_void_p),
    ]


class _NpState(C.Structure):
    _fields_ = [
        ("amount", C.c_void_p), ("amount_kind", C.c_void_p), ("stocks", C.c_void_p), ("cool", C.c_void_p),
        ("day", C.c_void_p), ("total", C.c_void_p), ("total_kind", C.c_void_p), ("gamma_reward", C.c_void_p),
        ("gr_kind", C.c_void_p), ("init_total", C.c_void_p), ("episode_return", C.c_void_p),
    ]


class NpTradingOracle:
    """N independent copies of the reference numpy ``StockTradingEnv``
    (finrl/meta/env_stock_trading/env_stocktrading_np.py), stepped on the CPU."""

    def __init__(self, price_array, tech_array, turbulence_array, n_envs, gamma=0.99, turbulence_thresh=99,
                 min_stock_rate=0.1, max_stock=1e2, initial_capital=1e6, buy_cost_pct=1e-3, sell_cost_pct=1e-3,
                 reward_scaling=2**-11, initial_stocks=None, obs_amount_floor=None):
        self.price, self.tech, self.turb_bool, self.turb_ary = np_tables(price_array, tech_array, turbulence_array, turbulence_thresh)
        T, D = self.price.shape
        N = int(n_envs)
        self.N, self.D, self.T, self.TD = N, D, T, self.tech.shape[1]
        self.O = 3 + 3 * D + self.TD
        self.init_stocks = np.ascontiguousarray(initial_stocks if initial_stocks is not None else np.zeros(D), dtype=np.float32)
        self.amount = np.zeros(N); self.amount_kind = np.zeros(N, dtype=np.uint8)
        self.stocks = np.zeros((N, D), dtype=np.float32); self.cool = np.zeros((N, D), dtype=np.float32)
        self.day = np.zeros(N, dtype=np.int32)
        self.total = np.zeros(N); self.total_kind = np.zeros(N, dtype=np.uint8)
        self.gamma_reward = np.zeros(N); self.gr_kind = np.zeros(N, dtype=np.uint8)
        self.init_total = np.zeros(N); self.episode_return = np.zeros(N)
        self._cfg = _NpCfg(N, D, self.TD, T, float(gamma), float(max_stock), float(min_stock_rate), float(buy_cost_pct),
                           float(sell_cost_pct), float(reward_scaling), float(initial_capital),
                           float("-inf") if obs_amount_floor is None else float(obs_amount_floor), _p(self.price),
                           _p(self.tech), _p(self.turb_bool), _p(self.turb_ary), _p(self.init_stocks))
        self._st = _NpState(_p(self.amount), _p(self.amount_kind), _p(self.stocks), _p(self.cool), _p(self.day),
                            _p(self.total), _p(self.total_kind), _p(self.gamma_reward), _p(self.gr_kind),
                            _p(self.init_total), _p(self.episode_return))
        self.reset()

    def obs(self):
        out = np.empty((self.N, self.O), dtype=np.float32)
        lib().ora_np_obs(C.byref(self._cfg), C.byref(self._st), _p(out))
        return out

    def reset(self, mask=None, stocks0=None, factor=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        s0 = None if stocks0 is None else np.ascontiguousarray(stocks0, dtype=np.float32).reshape(self.N, self.D)
        f = None if factor is None else np.ascontiguousarray(factor, dtype=np.float64).reshape(self.N)
        lib().ora_np_reset(C.byref(self._cfg), C.byref(self._st), _p(m), _p(s0), _p(f))
        return self.obs()

    def step(self, actions, want_obs=True):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        assert a.shape == (self.N, self.D)
        reward = np.empty(self.N); rk = np.empty(self.N, dtype=np.uint8)
        flags = np.empty(self.N, dtype=np.uint8)
        obs = np.empty((self.N, self.O), dtype=np.float32) if want_obs else None
        lib().ora_np_step(C.byref(self._cfg), C.byref(self._st), _p(a), _p(reward), _p(rk), _p(flags), _p(obs))
        return obs, reward, rk, flags


# --------------------------------------------------------------------------------------------
# A3
# --------------------------------------------------------------------------------------------
class _PfCfg(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("stock_dim", C.c_int32), ("n_tech", C.c_int32), ("n_days", C.c_int32),
                ("initial_amount", C.c_double), ("close", C.c_void_p), ("cov", C.c_void_p), ("tech", C.c_void_p)]


class _PfState(C.Structure):
    _fields_ = [("pv", C.c_void_p), ("day", C.c_void_p), ("reward", C.c_void_p)]


class PortfolioOracle:
    """N independent copies of the reference ``StockPortfolioEnv``
    (finrl/meta/env_portfolio_allocation/env_portfolio.py), stepped on the CPU."""

    def __init__(self, close, cov, tech, n_envs, initial_amount=1_000_000):
        self.close = np.ascontiguousarray(close, dtype=np.float64)
        T, D = self.close.shape
        self.cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(T, D, D)
        self.tech = np.ascontiguousarray(tech, dtype=np.float64).reshape(-1, T, D)
        K = self.tech.shape[0]
        N = int(n_envs)
        self.N, self.D, self.K, self.T = N, D, K, T
        self.pv = np.zeros(N); self.day = np.zeros(N, dtype=np.int32); self.reward = np.zeros(N)
        self._cfg = _PfCfg(N, D, K, T, float(initial_amount), _p(self.close), _p(self.cov), _p(self.tech))
        self._st = _PfState(_p(self.pv), _p(self.day), _p(self.reward))
        self.reset()

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().ora_portfolio_reset(C.byref(self._cfg), C.byref(self._st), _p(m))
        return self.obs()

    def obs(self):
        out = np.empty((self.N, self.D + self.K, self.D))
        lib().ora_portfolio_obs(C.byref(self._cfg), C.byref(self._st), _p(out))
        return out

    def step(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions)
        assert a.shape == (self.N, self.D) and a.dtype in (np.float32, np.float64)
        reward = np.empty(self.N); flags = np.empty(self.N, dtype=np.uint8)
        w = np.empty((self.N, self.D)); pr = np.empty(self.N)
        lib().ora_portfolio_step(C.byref(self._cfg), C.byref(self._st), _p(a), C.c_int(a.dtype == np.float64),
                                 _p(reward), _p(flags), _p(w), _p(pr), C.c_int(int(auto_reset)))
        return reward, flags, w, pr


# --------------------------------------------------------------------------------------------
# A4
# --------------------------------------------------------------------------------------------
class _CpCfg(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("stock_dim", C.c_int32), ("n_cols", C.c_int32), ("n_days", C.c_int32),
                ("buy_cost_pct", C.c_double), ("sell_cost_pct", C.c_double), ("hmax", C.c_double),
                ("discrete_actions", C.c_int32), ("shares_increment", C.c_int32), ("use_turbulence", C.c_int32),
                ("turbulence_threshold", C.c_double), ("initial_amount", C.c_double),
                ("cash_penalty_proportion", C.c_double), ("patient", C.c_int32),
                ("close", C.c_void_p), ("turb", C.c_void_p), ("info", C.c_void_p),
                ("hmax_vec", C.c_void_p), ("hmax_vec_f32", C.c_int32)]


def _hmax_args(hmax, D):
    """Scalar hmax -> (value, None, 0); per-asset array -> (0.0, float64 array [D], is_float32)."""
    if np.isscalar(hmax):
        return float(hmax), None, 0
    arr = np.asarray(hmax)
    if arr.shape != (D,):
        raise ValueError(f"hmax array must have shape ({D},)")
    return 0.0, np.ascontiguousarray(arr, dtype=np.float64), int(arr.dtype == np.float32)


class _CpState(C.Structure):
    _fields_ = [("cash", C.c_void_p), ("hold", C.c_void_p), ("date_index", C.c_void_p), ("start", C.c_void_p),
                ("fresh", C.c_void_p), ("last_cash", C.c_void_p), ("last_total", C.c_void_p), ("sum_trades", C.c_void_p)]


class CashPenaltyOracle:
    """N independent copies of the reference ``StockTradingEnvCashpenalty``
    (finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py), stepped on the CPU.
    ``info`` is [T, D, C]: the daily_information_cols of every asset (asset-major like get_date_vector)."""

    def __init__(self, close, info, turb, n_envs, buy_cost_pct=3e-3, sell_cost_pct=3e-3, hmax=10, discrete_actions=False,
                 shares_increment=1, turbulence_threshold=None, initial_amount=1e6, cash_penalty_proportion=0.1,
                 patient=False):
        self.close = np.ascontiguousarray(close, dtype=np.float64)
        T, D = self.close.shape
        self.info = np.ascontiguousarray(info, dtype=np.float64).reshape(T, -1)
        Cc = self.info.shape[1] // D
        self.turb = np.ascontiguousarray(turb if turb is not None else np.zeros(T), dtype=np.float64)
        N = int(n_envs)
        self.N, self.D, self.C, self.T, self.O = N, D, Cc, T, 1 + D + D * Cc
        self.cash = np.zeros(N); self.hold = np.zeros((N, D)); self.date_index = np.zeros(N, dtype=np.int32)
        self.start = np.zeros(N, dtype=np.int32); self.fresh = np.zeros(N, dtype=np.uint8)
        self.last_cash = np.zeros(N); self.last_total = np.zeros(N); self.sum_trades = np.zeros(N)
        hm, self._hmax_vec, hm32 = _hmax_args(hmax, D)
        self._cfg = _CpCfg(N, D, Cc, T, float(buy_cost_pct), float(sell_cost_pct), hm, int(discrete_actions),
                           int(shares_increment), int(turbulence_threshold is not None),
                           float(turbulence_threshold if turbulence_threshold is not None else 0.0), float(initial_amount),
                           float(cash_penalty_proportion), int(patient), _p(self.close), _p(self.turb), _p(self.info),
                           _p(self._hmax_vec), hm32)
        self._st = _CpState(_p(self.cash), _p(self.hold), _p(self.date_index), _p(self.start), _p(self.fresh),
                            _p(self.last_cash), _p(self.last_total), _p(self.sum_trades))
        self.reset()

    def reset(self, mask=None, start_points=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        sp = None if start_points is None else np.ascontiguousarray(start_points, dtype=np.int32)
        lib().ora_cp_reset(C.byref(self._cfg), C.byref(self._st), _p(m), _p(sp))
        return self.obs()

    def obs(self):
        out = np.empty((self.N, self.O))
        lib().ora_cp_obs(C.byref(self._cfg), C.byref(self._st), _p(out))
        return out

    def step(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions)
        assert a.shape == (self.N, self.D) and a.dtype in (np.float32, np.float64)
        reward = np.empty(self.N); flags = np.empty(self.N, dtype=np.uint8)
        lib().ora_cp_step(C.byref(self._cfg), C.byref(self._st), _p(a), C.c_int(a.dtype == np.float64), _p(reward),
                          _p(flags), C.c_int(int(auto_reset)))
        return reward, flags


# --------------------------------------------------------------------------------------------
# sibling: CryptoEnv
# --------------------------------------------------------------------------------------------
class _CryptoCfg(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("stock_dim", C.c_int32), ("tech_dim", C.c_int32), ("n_days", C.c_int32),
                ("lookback", C.c_int32), ("initial_capital", C.c_double), ("buy_cost_pct", C.c_double),
                ("sell_cost_pct", C.c_double), ("gamma", C.c_double), ("price", C.c_void_p), ("tech", C.c_void_p),
                ("act_norm", C.c_void_p)]


class _CryptoState(C.Structure):
    _fields_ = [("cash", C.c_void_p), ("stocks", C.c_void_p), ("time", C.c_void_p), ("total", C.c_void_p),
                ("gamma_return", C.c_void_p), ("episode_return", C.c_void_p)]


def crypto_action_norm(price_array):
    """_generate_action_normalizer (env_multiple_crypto.py:103-111)."""
    import math

    return np.asarray([1 / (10 ** math.floor(math.log(p, 10))) for p in price_array[0]]) * 10000


class CryptoOracle:
    """N independent copies of the reference ``CryptoEnv``
    (finrl/meta/env_cryptocurrency_trading/env_multiple_crypto.py), stepped on the CPU."""

    def __init__(self, price_array, tech_array, n_envs, lookback=1, initial_capital=1e6, buy_cost_pct=1e-3,
                 sell_cost_pct=1e-3, gamma=0.99):
        self.price = np.ascontiguousarray(price_array, dtype=np.float64)
        T, D = self.price.shape
        self.tech = np.ascontiguousarray(tech_array, dtype=np.float64).reshape(T, -1)
        self.norm = np.ascontiguousarray(crypto_action_norm(self.price), dtype=np.float64)
        N = int(n_envs)
        self.N, self.D, self.T, self.TD, self.lookback = N, D, T, self.tech.shape[1], int(lookback)
        self.O = 1 + D + self.TD * self.lookback
        self.cash = np.zeros(N); self.stocks = np.zeros((N, D), dtype=np.float32); self.time = np.zeros(N, dtype=np.int32)
        self.total = np.zeros(N); self.gamma_return = np.zeros(N); self.episode_return = np.zeros(N)
        self._cfg = _CryptoCfg(N, D, self.TD, T, self.lookback, float(initial_capital), float(buy_cost_pct),
                               float(sell_cost_pct), float(gamma), _p(self.price), _p(self.tech), _p(self.norm))
        self._st = _CryptoState(_p(self.cash), _p(self.stocks), _p(self.time), _p(self.total), _p(self.gamma_return),
                                _p(self.episode_return))
        self.reset()

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().ora_crypto_reset(C.byref(self._cfg), C.byref(self._st), _p(m))
        return self.obs()

    def obs(self):
        out = np.empty((self.N, self.O), dtype=np.float32)
        lib().ora_crypto_obs(C.byref(self._cfg), C.byref(self._st), _p(out))
        return out

    def step(self, actions, want_obs=True):
        a = np.ascontiguousarray(actions)
        assert a.shape == (self.N, self.D) and a.dtype in (np.float32, np.float64)
        reward = np.empty(self.N); flags = np.empty(self.N, dtype=np.uint8)
        obs = np.empty((self.N, self.O), dtype=np.float32) if want_obs else None
        lib().ora_crypto_step(C.byref(self._cfg), C.byref(self._st), _p(a), C.c_int(a.dtype == np.float64), _p(reward),
                              _p(flags), _p(obs))
        return obs, reward, flags


# --------------------------------------------------------------------------------------------
# sibling: StockTradingEnvStopLoss
# --------------------------------------------------------------------------------------------
class _SlCfg(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("stock_dim", C.c_int32), ("n_cols", C.c_int32), ("n_days", C.c_int32),
                ("buy_cost_pct", C.c_double), ("sell_cost_pct", C.c_double), ("hmax", C.c_double),
                ("discrete_actions", C.c_int32), ("shares_increment", C.c_int32),
                ("stoploss_penalty", C.c_double), ("profit_loss_ratio", C.c_double), ("use_turbulence", C.c_int32),
                ("turbulence_threshold", C.c_double), ("initial_amount", C.c_double),
                ("cash_penalty_proportion", C.c_double), ("patient", C.c_int32),
                ("close", C.c_void_p), ("turb", C.c_void_p), ("info", C.c_void_p),
                ("hmax_vec", C.c_void_p), ("hmax_vec_f32", C.c_int32)]


class _SlState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("cash", "hold", "prev_hold", "avg_buy", "n_buys", "cdiff", "pdiff", "date_index",
                                          "start", "fresh", "last_cash", "last_total", "sum_trades")]


class StopLossOracle:
    """N independent copies of the reference ``StockTradingEnvStopLoss``
    (finrl/meta/env_stock_trading/env_stocktrading_stoploss.py), stepped on the CPU."""

    def __init__(self, close, info, turb, n_envs, buy_cost_pct=3e-3, sell_cost_pct=3e-3, hmax=10, discrete_actions=False,
                 shares_increment=1, stoploss_penalty=0.9, profit_loss_ratio=2, turbulence_threshold=None,
                 initial_amount=1e6, cash_penalty_proportion=0.1, patient=False):
        self.close = np.ascontiguousarray(close, dtype=np.float64)
        T, D = self.close.shape
        self.info = np.ascontiguousarray(info, dtype=np.float64).reshape(T, -1)
        Cc = self.info.shape[1] // D
        self.turb = np.ascontiguousarray(turb if turb is not None else np.zeros(T), dtype=np.float64)
        N = int(n_envs)
        self.N, self.D, self.C, self.T, self.O = N, D, Cc, T, 1 + D + D * Cc
        z = lambda: np.zeros((N, D))
        self.cash = np.zeros(N); self.hold = z(); self.prev_hold = z(); self.avg_buy = z(); self.n_buys = z()
        self.cdiff = z(); self.pdiff = z()
        self.date_index = np.zeros(N, dtype=np.int32); self.start = np.zeros(N, dtype=np.int32)
        self.fresh = np.zeros(N, dtype=np.uint8)
        self.last_cash = np.zeros(N); self.last_total = np.zeros(N); self.sum_trades = np.zeros(N)
        hm, self._hmax_vec, hm32 = _hmax_args(hmax, D)
        self._cfg = _SlCfg(N, D, Cc, T, float(buy_cost_pct), float(sell_cost_pct), hm, int(discrete_actions),
                           int(shares_increment), float(stoploss_penalty), float(profit_loss_ratio),
                           int(turbulence_threshold is not None),
                           float(turbulence_threshold if turbulence_threshold is not None else 0.0), float(initial_amount),
                           float(cash_penalty_proportion), int(patient), _p(self.close), _p(self.turb), _p(self.info),
                           _p(self._hmax_vec), hm32)
        self._st = _SlState(_p(self.cash), _p(self.hold), _p(self.prev_hold), _p(self.avg_buy), _p(self.n_buys),
                            _p(self.cdiff), _p(self.pdiff), _p(self.date_index), _p(self.start), _p(self.fresh),
                            _p(self.last_cash), _p(self.last_total), _p(self.sum_trades))
        self.reset()

    def reset(self, mask=None, start_points=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        sp = None if start_points is None else np.ascontiguousarray(start_points, dtype=np.int32)
        lib().ora_sl_reset(C.byref(self._cfg), C.byref(self._st), _p(m), _p(sp))
        return self.obs()

    def obs(self):
        out = np.empty((self.N, self.O))
        lib().ora_sl_obs(C.byref(self._cfg), C.byref(self._st), _p(out))
        return out

    def step(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions)
        assert a.shape == (self.N, self.D) and a.dtype in (np.float32, np.float64)
        reward = np.empty(self.N); flags = np.empty(self.N, dtype=np.uint8)
        lib().ora_sl_step(C.byref(self._cfg), C.byref(self._st), _p(a), C.c_int(a.dtype == np.float64), _p(reward),
                          _p(flags), C.c_int(int(auto_reset)))
        return reward, flags
