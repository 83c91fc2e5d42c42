"""Shared host plumbing of the batched engines: device binding, action marshalling, and the two
C-ABI call shapes (``*_step`` and ``*_rollout``) every env kind exposes with identical signatures."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


class BatchedEnvBase:
    """Subclasses set ``_PREFIX`` (e.g. ``"frl_trading"``), build ``self._p`` (the params struct) and the
    buffers ``self._obs`` / ``self._flags`` / ``self.stats`` and, unless the reward lives in the state
    (StockTradingEnv), ``self._rew``."""

    _PREFIX = ""
    _ACTION_NAME = "stock_dim"

    # ---- construction helpers ---------------------------------------------------------------
    def _bind_device(self, device):
        import torch

        self._torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.EngineError("finrl_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _cabi.lib()  # fail loudly before any allocation if the extension is missing
        self.launches = 0          # kernels launched through the C-ABI (bench.py reports it)
        self.kernel_events = None  # set to a list to collect (start, end) CUDA events around each step launch
        return torch

    def _stream(self):
        return _cabi.current_stream(self.device)

    def seed(self, seed=None):
        """Seed of the in-kernel reset draws (numpy env ``if_train``, cash-penalty / stop-loss ``random_start``).
        Unseeded engines take 64 bits from the OS, as the reference's envs seed themselves from the clock."""
        import os

        self._seed0 = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2**64 - 1)
        self._launch_id = 0
        return [seed]

    def _next_reset_seed(self):
        """A fresh stream id per launch for params structs that carry ``reset_seed`` (no-op for the others)."""
        p = self._p
        if hasattr(p, "reset_seed"):
            if not hasattr(self, "_seed0"):
                self.seed(None)
            self._launch_id += 1
            p.reset_seed = (self._seed0 + self._launch_id * 0x9E3779B97F4A7C15) & (2**64 - 1)

    def _fn(self, name):
        return getattr(_cabi.lib(), f"{self._PREFIX}_{name}")

    def _as_actions(self, actions, ndim):
        torch = self._torch
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype not in (torch.float32, torch.float64):
            actions = actions.to(torch.float32)
        if actions.device != self.device:
            actions = actions.to(self.device, non_blocking=True)
        if actions.dim() != ndim or actions.shape[-1] != self.stock_dim:
            raise ValueError(f"actions must have {ndim} dims ending in {self._ACTION_NAME}={self.stock_dim}, got {tuple(actions.shape)}")
        return actions.contiguous()

    def _obs_out(self):
        return self._obs

    def _reward_out(self):
        return getattr(self, "_rew", None)

    def _reward_result(self):
        return self._rew

    # ---- the two call shapes ----------------------------------------------------------------------
    def step(self, actions, auto_reset: bool = False, want_obs: bool = True, accumulate_stats: bool = False,
             want_done: bool = True):
        """One ``step`` of every env.  Returns (obs, reward[N] f64, done[N] bool, flags[N] u8).  The tensors
        are engine-owned buffers that the next call overwrites.  ``done`` is ``flags & FLAG_DONE`` (two tiny
        torch kernels); pass ``want_done=False`` to get None instead."""
        torch = self._torch
        a = self._as_actions(actions, 2)
        if a.shape[0] != self.n_envs:
            raise ValueError(f"actions must have n_envs={self.n_envs} rows")
        obs = self._obs_out() if want_obs else None
        self._next_reset_seed()
        ev = self.kernel_events
        with torch.cuda.device(self.device):
            if ev is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = self._fn("step")(
                C.byref(self._p), _cabi.ptr(a), int(a.dtype == torch.float64), _cabi.ptr(self._reward_out()),
                _cabi.ptr(self._flags), _cabi.ptr(obs), int(auto_reset),
                self._stats_arg() if accumulate_stats else None, self._stream(),
            )
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
        _cabi.check(rc, f"{self._PREFIX}_step")
        self.launches += 1
        done = (self._flags & _cabi.FLAG_DONE).bool() if want_done else None
        return self._shape_obs(obs), self._reward_result(), done, self._flags

    def _shape_obs(self, obs):
        return obs

    def rollout(self, actions, layout: str = "KND", obs_mode: str = "last", auto_reset: bool = True,
                accumulate_stats: bool = True, rewards=None, flags=None, obs=None):
        """Fused multi-step rollout: ``actions`` is [K, N, D] (layout "KND", time-major like SB3's rollout
        buffer) or [N, K, D] ("NKD").  Returns (obs, rewards[K,N] f64, flags[K,N] u8) where obs is None /
        [N,O] / [K,N,O] for obs_mode "none" / "last" / "all"."""
        torch = self._torch
        a = self._as_actions(actions, 3)
        D, N = self.stock_dim, self.n_envs
        if layout == "KND":
            K, ok = a.shape[0], a.shape[1] == N
        elif layout == "NKD":
            K, ok = a.shape[1], a.shape[0] == N
        else:
            raise ValueError("layout must be 'KND' or 'NKD'")
        if not ok:
            raise ValueError(f"actions shape {tuple(a.shape)} does not match n_envs={N} for layout {layout}")
        step_stride, env_stride = (N * D, D) if layout == "KND" else (D, K * D)
        mode = {"none": _cabi.OBS_NONE, "last": _cabi.OBS_LAST, "all": _cabi.OBS_ALL}[obs_mode]
        if rewards is None:
            rewards = torch.empty((K, N), dtype=torch.float64, device=self.device)
        if flags is None:
            flags = torch.empty((K, N), dtype=torch.uint8, device=self.device)
        if mode == _cabi.OBS_LAST and obs is None:
            obs = self._obs_out()
        elif mode == _cabi.OBS_ALL and obs is None:
            obs = torch.empty((K, N, self._obs_out().shape[-1]), dtype=torch.float32, device=self.device)
        self._next_reset_seed()
        with torch.cuda.device(self.device):
            _cabi.check(
                self._fn("rollout")(
                    C.byref(self._p), _cabi.ptr(a), int(a.dtype == torch.float64), step_stride, env_stride, int(K),
                    _cabi.ptr(rewards), _cabi.ptr(flags), _cabi.ptr(obs) if mode else None, mode, int(auto_reset),
                    self._stats_arg() if accumulate_stats else None, self._stream(),
                ),
                f"{self._PREFIX}_rollout",
            )
        self.launches += 1
        return (obs if mode else None), rewards, flags

    def observe(self, out=None):
        """float32 observation of every env, [N, O]."""
        out = self._obs_out() if out is None else out
        with self._torch.cuda.device(self.device):
            _cabi.check(self._fn("observe")(C.byref(self._p), _cabi.ptr(out), self._stream()), f"{self._PREFIX}_observe")
        self.launches += 1
        return self._shape_obs(out)

    def clone_state_into(self, other):
        """Copy every per-env state array of this engine into ``other`` (an engine of the same class and size
        built on the same tables) — the device half of ``copy.deepcopy(env)`` that the reference's
        ``get_sb_env`` performs (env_stocktrading_cashpenalty.py:374-380)."""
        torch = self._torch
        if type(other) is not type(self) or other.n_envs != self.n_envs:
            raise ValueError("clone_state_into needs an engine of the same class and n_envs")
        for name, v in vars(self).items():
            if name in ("stats", "_stats_block", "_stats_total"):
                continue  # statistics (and a peer binding) belong to the engine they were attached to
            w = vars(other).get(name)
            if isinstance(v, torch.Tensor) and isinstance(w, torch.Tensor) and w.shape == v.shape \
                    and w.dtype == v.dtype and w.data_ptr() != v.data_ptr():
                w.copy_(v)
        return other

    def _mask(self, mask):
        if mask is None:
            return None
        m = self._torch.as_tensor(mask, device=self.device).to(self._torch.uint8).contiguous()
        if m.shape != (self.n_envs,):
            raise ValueError("mask must have shape [n_envs]")
        return m

    _STAT7_NAME = _cabi.STAT_NAMES[7]

    def use_stats_block(self, block, alternate: bool = False):
        """Accumulate into ``block`` (a 48-double frl_stats_block on this device) instead of the engine's own;
        :class:`finrl_b200.dist.StatsExchange` hands every rank's engine a peer-mapped block this way, with
        ``alternate=True``: consecutive launches then use the block's two accumulators in turn and each launch
        pushes the previous one's sums to the peers (frl_stats_block in the header)."""
        if block.numel() != _cabi.STATS_BLOCK_DOUBLES or block.dtype != self._torch.float64 or not block.is_cuda \
                or block.data_ptr() % 128:
            raise ValueError("use_stats_block needs a 128-byte-aligned float64 CUDA tensor of 48 elements")
        self._stats_block = block
        self.stats = block[: _cabi.N_STATS]
        self._stats_alternate = alternate
        self._stats_turn = 0

    def _stats_arg(self):
        """The accumulator this launch adds into: sum[0], or sum[0] / sum[1] in turn under an exchange."""
        if not getattr(self, "_stats_alternate", False):
            return _cabi.ptr(self.stats)
        t = self._stats_turn
        self._stats_turn = t ^ 1
        return C.c_void_p(self._stats_block.data_ptr() + 64 * t)

    def read_stats(self, reset: bool = False):
        """The statistics accumulated in this engine's block as a dict: both accumulators plus, under an exchange,
        the totals that already arrived (all ranks' launches); ``StatsExchange.totals()`` is the synchronised read."""
        n = _cabi.N_STATS
        blk = self._stats_block
        vals = (blk[:n] + blk[n : 2 * n] + blk[2 * n : 3 * n]).tolist()
        if reset:
            blk[: 3 * n].zero_()
        names = list(_cabi.STAT_NAMES)
        names[7] = self._STAT7_NAME
        return dict(zip(names, vals))
