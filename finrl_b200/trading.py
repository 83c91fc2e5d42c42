"""Batched StockTradingEnv: N independent copies of the reference env stepped by one CUDA kernel.

Reference: /root/reference/finrl/meta/env_stock_trading/env_stocktrading.py (``StockTradingEnv``).
Same constructor keywords, same ``reset`` / ``step`` semantics (incl. quirks Q1-Q5 of SURVEY.md
§8a), but every method acts on all N envs and returns device tensors.  The arithmetic happens in
``finrl_b200/csrc/trading.cu`` behind the C-ABI of ``include/finrl_b200.h``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from ._base import BatchedEnvBase
from .tables import TradingTables


def _scalar_cost(x, name):
    """The reference only works with scalar costs (``1 + list`` raises, quirk Q7); its tutorials pass
    ``[c] * D``.  Accept a scalar or a constant vector."""
    if np.isscalar(x):
        return float(x)
    arr = np.asarray(x, dtype=np.float64).reshape(-1)
    if arr.size == 0 or not np.all(arr == arr[0]):
        raise NotImplementedError(f"{name}: per-stock costs that differ are not supported (the reference cannot consume them either)")
    return float(arr[0])


class FactoredObs:
    """The observation batch of a host-resident caller in factored form: ``env_part[N, 1+D]`` float32 (cash,
    holdings), ``state_day[N]`` int32 and the per-day template ``tmpl[T, O]`` the host holds once.  Row ``n`` of
    the dense observation is ``tmpl[state_day[n]]`` with the cash / holdings slots taken from ``env_part[n]`` —
    bit-identical to what ``step()`` writes.  ``fo[n]`` builds one row, ``fo.dense()`` all of them (native,
    multi-threaded); 137 instead of 1213 bytes per DOW-30 env cross PCIe."""

    def __init__(self, env_part, state_day, tmpl, stock_dim):
        self.env_part, self.state_day, self.tmpl, self.stock_dim = env_part, state_day, tmpl, int(stock_dim)

    def __len__(self):
        return int(self.env_part.shape[0])

    @property
    def shape(self):
        return (len(self), int(self.tmpl.shape[1]))

    def __getitem__(self, n):
        D = self.stock_dim
        e = self.env_part[n].numpy() if hasattr(self.env_part, "numpy") else np.asarray(self.env_part[n])
        row = np.array(self.tmpl[int(self.state_day[n])], dtype=np.float32)
        row[0] = e[0]
        row[1 + D : 1 + 2 * D] = e[1:]
        return row

    def dense(self, out=None, n_threads: int = 0):
        """Dense [N, O] float32 array (written into ``out`` when given) via frl_expand_obs_host."""
        N, O = self.shape
        ep = self.env_part.numpy() if hasattr(self.env_part, "numpy") else np.ascontiguousarray(self.env_part, dtype=np.float32)
        sd = self.state_day.numpy() if hasattr(self.state_day, "numpy") else np.ascontiguousarray(self.state_day, dtype=np.int32)
        if out is None:
            out = np.empty((N, O), dtype=np.float32)
        dst = out.numpy() if hasattr(out, "numpy") else out
        if dst.shape != (N, O) or dst.dtype != np.float32 or not dst.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float32 array of shape {(N, O)}")
        tm = self.tmpl
        _cabi.check(_cabi.lib().frl_expand_obs_host(
            tm.ctypes.data, tm.shape[0], O, self.stock_dim, ep.ctypes.data, sd.ctypes.data, N, dst.ctypes.data, int(n_threads)),
            "frl_expand_obs_host")
        return out


class BatchedStockTradingEnv(BatchedEnvBase):
    """N lock-stepped (or not) ``StockTradingEnv`` instances on one GPU.

    Parameters mirror ``StockTradingEnv.__init__`` (env_stocktrading.py:24-47); extra keywords:
    ``n_envs``, ``device`` and ``tables`` (pre-built :class:`TradingTables`, instead of ``df``).
    ``step`` / ``rollout`` / ``observe`` / ``read_stats`` come from :class:`BatchedEnvBase`.
    """

    _PREFIX = "frl_trading"

    def __init__(
        self,
        df=None,
        stock_dim: int = None,
        hmax: float = 100,
        initial_amount: float = 1_000_000,
        num_stock_shares: Optional[Sequence[int]] = None,
        buy_cost_pct=0.001,
        sell_cost_pct=0.001,
        reward_scaling: float = 1e-4,
        state_space: Optional[int] = None,
        action_space: Optional[int] = None,
        tech_indicator_list: Sequence[str] = (),
        turbulence_threshold=None,
        risk_indicator_col="turbulence",
        day: int = 0,
        initial: bool = True,
        previous_state: Sequence[float] = (),
        *,
        n_envs: int = 1,
        device="cuda",
        tables: Optional[TradingTables] = None,
        track_asset: bool = False,
    ):
        torch = self._bind_device(device)
        if tables is None:
            if df is None:
                raise ValueError("either df or tables is required")
            if turbulence_threshold is not None and risk_indicator_col not in df.columns:
                raise KeyError(risk_indicator_col)
            tables = TradingTables.from_frame(df, stock_dim, list(tech_indicator_list), risk_indicator_col, self.device)
        self.tables = tables
        self.df = df
        D, K, T = tables.stock_dim, tables.n_tech, tables.n_days
        if stock_dim is not None and int(stock_dim) != D:
            raise ValueError(f"stock_dim={stock_dim} but the tables hold {D} stocks")
        O = tables.obs_dim
        if state_space is not None and int(state_space) != O:
            raise ValueError(f"state_space={state_space} != 1 + 2*stock_dim + len(tech)*stock_dim = {O}")
        if action_space is not None and int(action_space) != D:
            raise ValueError(f"action_space={action_space} != stock_dim={D}")
        self.n_envs, self.stock_dim, self.n_tech, self.n_days = int(n_envs), D, K, T
        self.state_space = O
        self.hmax = hmax
        self.initial_amount = initial_amount
        self.buy_cost_pct = _scalar_cost(buy_cost_pct, "buy_cost_pct")
        self.sell_cost_pct = _scalar_cost(sell_cost_pct, "sell_cost_pct")
        self.reward_scaling = reward_scaling
        self.turbulence_threshold = turbulence_threshold
        self.risk_indicator_col = risk_indicator_col
        self.tech_indicator_list = list(tech_indicator_list)
        self.initial = initial
        self.previous_state = list(previous_state)
        if int(hmax) * (T + 1) >= 2**31:
            raise ValueError("hmax * n_days must stay below 2^31 (holdings are int32)")
        if not 0 <= int(day) < T:
            raise ValueError(f"day={day} outside [0, {T})")

        if initial:
            init_cash = float(initial_amount)
            shares = [0] * D if num_stock_shares is None else list(num_stock_shares)
        else:  # resume from a previous run's state list (env_stocktrading.py:423-450)
            if len(self.previous_state) < 1 + 2 * D:
                raise ValueError("previous_state must hold [cash, prices x D, holdings x D, ...]")
            init_cash = float(self.previous_state[0])
            shares = [int(round(float(v))) for v in self.previous_state[D + 1 : 2 * D + 1]]
        if len(shares) != D:
            raise ValueError(f"num_stock_shares must have {D} entries")
        self.num_stock_shares = shares
        self._init_cash = init_cash

        N = self.n_envs
        dev = self.device
        self.init_hold = torch.tensor(shares, dtype=torch.int32, device=dev)
        self.cash = torch.empty(N, dtype=torch.float64, device=dev)
        self.hold = torch.empty((D, N), dtype=torch.int32, device=dev)  # stock-major
        self.day = torch.empty(N, dtype=torch.int32, device=dev)
        self.sday = torch.empty(N, dtype=torch.int32, device=dev)
        self.cost = torch.empty(N, dtype=torch.float64, device=dev)
        self.trades = torch.empty(N, dtype=torch.int32, device=dev)
        self.reward = torch.empty(N, dtype=torch.float64, device=dev)
        self.episode = torch.empty(N, dtype=torch.int32, device=dev)
        self._stats_block = _cabi.new_stats_block(torch, dev)
        self.stats = self._stats_block[:_cabi.N_STATS]
        self._obs = torch.empty((N, O), dtype=torch.float32, device=dev)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)

        p = _cabi.TradingParams()
        p.n_envs, p.stock_dim, p.n_tech, p.n_days, p.obs_dim, p.env_stride = N, D, K, T, O, N
        p.hmax = float(hmax)
        p.initial_amount = init_cash
        p.buy_cost_pct, p.sell_cost_pct = self.buy_cost_pct, self.sell_cost_pct
        p.reward_scaling = float(reward_scaling)
        p.use_turbulence = int(turbulence_threshold is not None)
        p.close_pitch = int(tables.close.shape[1])
        p.turbulence_threshold = float(turbulence_threshold) if turbulence_threshold is not None else 0.0
        p.close, p.disable_mask = tables.close.data_ptr(), tables.disable_mask.data_ptr()
        p.risk, p.obs_tmpl = tables.risk.data_ptr(), tables.obs_tmpl.data_ptr()
        p.init_hold = self.init_hold.data_ptr()
        p.cash, p.hold, p.day, p.sday = self.cash.data_ptr(), self.hold.data_ptr(), self.day.data_ptr(), self.sday.data_ptr()
        p.cost, p.trades = self.cost.data_ptr(), self.trades.data_ptr()
        p.reward, p.episode = self.reward.data_ptr(), self.episode.data_ptr()
        # optional: the reference's end_total_asset of every env after each launch (asset_memory)
        self.asset = torch.zeros(N, dtype=torch.float64, device=dev) if track_asset else None
        p.asset_out = self.asset.data_ptr() if track_asset else None
        tmpl4 = getattr(tables, "obs_tmpl4", None)
        p.obs_tmpl4 = tmpl4.data_ptr() if tmpl4 is not None else None
        self._p = p
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().frl_trading_init(C.byref(p), int(day), self._stream()), "frl_trading_init")
        self.launches += 1

    # ------------------------------------------------------------------------------------------
    def _reward_out(self):
        return None  # the reward lives in the state (self.reward), like the reference's self.reward

    def _reward_result(self):
        return self.reward

    def reset(self, mask=None, out=None):
        """``StockTradingEnv.reset`` for all envs (or those with ``mask[n] != 0``)."""
        out = self._obs if out is None else out
        mask = self._mask(mask)
        with self._torch.cuda.device(self.device):
            _cabi.check(self._fn("reset")(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(out), self._stream()), "frl_trading_reset")
        self.launches += 2
        return out

    # ------------------------------------------------------------------------------------------
    def _chunk_params(self, start: int, count: int):
        """A params struct addressing envs [start, start+count): same tables, state pointers offset,
        stock-major arrays keep the full leading dimension (that is what env_stride is for)."""
        q = _cabi.TradingParams()
        C.memmove(C.byref(q), C.byref(self._p), C.sizeof(q))
        q.n_envs = count
        q.cash = self.cash.data_ptr() + 8 * start
        q.hold = self.hold.data_ptr() + 4 * start
        q.day = self.day.data_ptr() + 4 * start
        q.sday = self.sday.data_ptr() + 4 * start
        q.cost = self.cost.data_ptr() + 8 * start
        q.trades = self.trades.data_ptr() + 4 * start
        q.reward = self.reward.data_ptr() + 8 * start
        q.episode = self.episode.data_ptr() + 4 * start
        q.asset_out = (self.asset.data_ptr() + 8 * start) if self.asset is not None else None
        return q

    def observe_factored(self, env_part=None, state_day=None):
        """The current observation in factored form (see :class:`FactoredObs`): device tensors
        ``env_part[N, 1+D]`` float32 and ``state_day[N]`` int32."""
        torch = self._torch
        if env_part is None:
            env_part = self._factored_bufs()[0]
        if state_day is None:
            state_day = self._factored_bufs()[1]
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().frl_trading_observe_factored(C.byref(self._p), _cabi.ptr(env_part), _cabi.ptr(state_day),
                                                                 self._stream()), "frl_trading_observe_factored")
        self.launches += 1
        return env_part, state_day

    def _factored_bufs(self):
        torch = self._torch
        if getattr(self, "_fact", None) is None:
            self._fact = (torch.empty((self.n_envs, 1 + self.stock_dim), dtype=torch.float32, device=self.device),
                          torch.empty(self.n_envs, dtype=torch.int32, device=self.device))
        return self._fact

    def make_host_buffers(self, obs_layout: str = "dense", action_dtype=None):
        """Pinned host buffers for :meth:`step_host`: (actions, obs, reward, flags); ``obs`` is a float32 [N, O]
        tensor for ``"dense"`` and a :class:`FactoredObs` over pinned tensors for ``"factored"``."""
        torch = self._torch
        N, D, O = self.n_envs, self.stock_dim, self.state_space
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()  # noqa: E731
        act = pin((N, D), action_dtype or torch.float32)
        if obs_layout == "dense":
            obs = pin((N, O), torch.float32)
        elif obs_layout == "factored":
            obs = FactoredObs(pin((N, 1 + D), torch.float32), pin((N,), torch.int32), self.tables.host_tmpl, D)
        else:
            raise ValueError("obs_layout must be 'dense' or 'factored'")
        return act, obs, pin((N,), torch.float64), pin((N,), torch.uint8)

    def _expand_threads(self):
        """Host threads for the dense-by-expansion path: this process's share of the CPUs it may run on
        (``FRL_EXPAND_THREADS`` overrides; 14 / 16 / 20 threads and pieces of 512 / 2048 / 8192 rows all measure the same)."""
        import os

        try:
            cpus = len(os.sched_getaffinity(0))
        except AttributeError:  # pragma: no cover
            cpus = os.cpu_count() or 1
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
        if os.environ.get("FRL_EXPAND_THREADS"):
            return max(1, int(os.environ["FRL_EXPAND_THREADS"]))
        return max(1, min(16, cpus // max(local_world, 1)))

    def step_host(self, actions_host, obs_host, reward_host, flags_host, auto_reset: bool = True, n_chunks: int = 8,
                  host_expand=None):
        """Host-resident agents: one ``step`` with pinned HOST buffers in and out, software-pipelined.

        The env range is cut into ``n_chunks`` slices; slice c's action upload (H2D), its kernel and its
        observation / reward / flag download (D2H) run on one of three streams, so the PCIe link stays busy
        in both directions while kernels run.  Returns after everything has landed in the host buffers.

        ``obs_host`` is either a dense float32 [N, O] tensor or a :class:`FactoredObs` (``make_host_buffers``):
        with the factored layout the step kernel writes no observation at all, a small kernel extracts the
        env-specific slots, and only 4*(1+D) + 4 bytes per env come back instead of 4*O (91 % of a DOW-30
        observation is the per-day row every env of that day shares).

        ``host_expand`` (dense ``obs_host`` only; default: on when this process has at least four CPUs to itself):
        the observation still crosses PCIe in factored form and the dense rows are rebuilt in ``obs_host`` by host
        threads (``frl_expand_obs_host_chunks``, non-temporal stores), slice by slice as the transfers land — the
        caller gets the same bits in the same buffer, PCIe carries a ninth of the bytes."""
        torch = self._torch
        N, D, O = self.n_envs, self.stock_dim, self.state_space
        factored = isinstance(obs_host, FactoredObs)
        dense_out = None
        if not factored:
            if host_expand is None:
                host_expand = self._expand_threads() >= 4 and self.tables.host_tmpl is not None
            if host_expand:
                if obs_host.shape != (N, O) or obs_host.dtype != torch.float32 or not obs_host.is_contiguous():
                    raise ValueError("step_host: bad buffer shapes")
                if getattr(self, "_fact_host", None) is None:
                    self._fact_host = self.make_host_buffers("factored")[1]
                dense_out, obs_host, factored = obs_host, self._fact_host, True
        obs_bufs = (obs_host.env_part, obs_host.state_day) if factored else (obs_host,)
        if not (actions_host.is_pinned() and reward_host.is_pinned() and flags_host.is_pinned() and all(b.is_pinned() for b in obs_bufs)):
            raise ValueError("step_host needs pinned host tensors")
        ok = obs_bufs[0].shape == ((N, 1 + D) if factored else (N, O)) and (not factored or obs_bufs[1].shape == (N,))
        if actions_host.shape != (N, D) or not ok or actions_host.dtype not in (torch.float32, torch.float64):
            raise ValueError("step_host: bad buffer shapes")
        if factored:
            d_env, d_sd = self._factored_bufs()
        if getattr(self, "_host_streams", None) is None:
            self._host_streams = [torch.cuda.Stream(self.device) for _ in range(3)]
            self._host_act = torch.empty((N, D), dtype=actions_host.dtype, device=self.device)
        if self._host_act.dtype != actions_host.dtype:
            self._host_act = torch.empty((N, D), dtype=actions_host.dtype, device=self.device)
        per = -(-N // n_chunks)
        per = -(-per // 32) * 32  # keep slices warp-tile aligned
        lib = _cabi.lib()
        f64 = int(actions_host.dtype == torch.float64)
        cur = torch.cuda.current_stream(self.device)
        for st in self._host_streams:
            st.wait_stream(cur)
        rc = 0
        landed = []  # (start, count, event) per slice, for the host-side expansion
        with torch.cuda.device(self.device):
            for c, start in enumerate(range(0, N, per)):
                cnt = min(per, N - start)
                st = self._host_streams[c % len(self._host_streams)]
                with torch.cuda.stream(st):
                    d_act = self._host_act[start : start + cnt]
                    d_act.copy_(actions_host[start : start + cnt], non_blocking=True)
                    q = self._chunk_params(start, cnt)
                    rc |= lib.frl_trading_step(C.byref(q), _cabi.ptr(d_act), f64, None, C.c_void_p(self._flags.data_ptr() + start),
                                               None if factored else C.c_void_p(self._obs.data_ptr() + 4 * O * start),
                                               int(auto_reset), None, C.c_void_p(st.cuda_stream))
                    if factored:
                        rc |= lib.frl_trading_observe_factored(
                            C.byref(q), C.c_void_p(d_env.data_ptr() + 4 * (1 + D) * start), C.c_void_p(d_sd.data_ptr() + 4 * start),
                            C.c_void_p(st.cuda_stream))
                        self.launches += 1
                        obs_host.env_part[start : start + cnt].copy_(d_env[start : start + cnt], non_blocking=True)
                        obs_host.state_day[start : start + cnt].copy_(d_sd[start : start + cnt], non_blocking=True)
                        if dense_out is not None:
                            ev = torch.cuda.Event()
                            ev.record(st)
                            landed.append((start, cnt, ev))
                    else:
                        obs_host[start : start + cnt].copy_(self._obs[start : start + cnt], non_blocking=True)
                    reward_host[start : start + cnt].copy_(self.reward[start : start + cnt], non_blocking=True)
                    flags_host[start : start + cnt].copy_(self._flags[start : start + cnt], non_blocking=True)
                self.launches += 1
        _cabi.check(rc, "frl_trading_step (step_host)")
        if dense_out is not None:
            # host threads rebuild the dense rows slice by slice, each slice as soon as its event has fired
            n = len(landed)
            starts = (C.c_int64 * n)(*[a for a, _, _ in landed])
            counts = (C.c_int64 * n)(*[b for _, b, _ in landed])
            events = (C.c_void_p * n)(*[e.cuda_event for _, _, e in landed])
            tm = self.tables.host_tmpl
            _cabi.check(lib.frl_expand_obs_host_chunks(
                tm.ctypes.data, tm.shape[0], O, D, obs_host.env_part.data_ptr(), obs_host.state_day.data_ptr(),
                dense_out.data_ptr(), n, starts, counts, events, self._expand_threads()), "frl_expand_obs_host_chunks")
        for st in self._host_streams:
            cur.wait_stream(st)
        cur.synchronize()

    def get_state(self):
        """Per-env state in the natural [N, ...] layout (host-friendly; holdings transposed)."""
        return {
            "cash": self.cash.clone(), "hold": self.hold.t().contiguous(), "day": self.day.clone(),
            "sday": self.sday.clone(), "cost": self.cost.clone(), "trades": self.trades.clone(),
            "reward": self.reward.clone(), "episode": self.episode.clone(),
        }

    def set_state(self, **kw):
        """Overwrite state arrays (``cash``, ``hold`` [N,D], ``day``, ``sday``, ``cost``, ``trades``,
        ``reward``, ``episode``) — the batched form of the reference's ``previous_state`` resume."""
        torch = self._torch
        for k, v in kw.items():
            dst = getattr(self, k)
            v = torch.as_tensor(v, device=self.device)
            if k == "hold":
                v = v.to(torch.int32).reshape(self.n_envs, self.stock_dim).t()
            dst.copy_(v.to(dst.dtype).reshape(dst.shape) if k != "hold" else v)

    def total_asset(self):
        """cash + sum(price * holdings) with the prices currently in the state list, in the reference's order
        (``state[0] + sum(prices * holdings)``: a sequential sum from 0, then the cash; fp64, exact per env)."""
        torch = self._torch
        sd = torch.where(self.sday < 0, -self.sday - 1, self.sday).long()
        prices = self.tables.close[sd, : self.stock_dim]  # [N, D]
        acc = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)
        for j in range(self.stock_dim):
            acc = acc + prices[:, j] * self.hold[j].double()
        return self.cash + acc
