"""Batched VecEnv: the SB3 ``VecEnv`` protocol (and ElegantRL's vectorised-env convention) over N GPU envs.

What the reference's agent adapters call:
  * SB3 (finrl/agents/stablebaselines3/models.py): ``num_envs``, ``observation_space``, ``action_space``,
    ``reset() -> obs[N,O] float32``, ``step(actions) -> (obs, rewards, dones, infos)`` with auto-reset on
    done and ``infos[i]["terminal_observation"]``, ``env_method``, ``get_attr``, ``render``, ``close``.
  * ElegantRL: ``env_num``, ``state_dim``, ``action_dim``, ``max_step``, ``if_discrete``, ``target_return``,
    ``reset() -> Tensor[env_num, state_dim]``, ``step(Tensor) -> (state, reward, done, info)`` tensors.

``BatchedVecEnv`` wraps any of the four batched engines.  numpy mode (default) returns host arrays like
``DummyVecEnv``; ``tensor_mode=True`` keeps everything on the device and lets the kernel auto-reset, which
is the path meant for 1M-env training loops (a Python list of a million info dicts is not).
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .spaces import Box, vec_env_base

_Base = vec_env_base()


class _LazyInfos:
    """Sequence of N info dicts built on demand (only done envs carry a terminal_observation)."""

    def __init__(self, n, terminal):
        self._n, self._terminal = n, terminal

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        t = self._terminal.get(i)
        return {} if t is None else {"terminal_observation": t}

    def __iter__(self):
        return (self[i] for i in range(self._n))


class BatchedVecEnv(_Base):
    def __init__(self, engine, action_low=-1.0, action_high=1.0, tensor_mode: bool = False, obs_shape=None,
                 copy_outputs: bool = True, record_memory: bool = False, dates=None, tickers=None):
        self.engine = engine
        # record_memory: keep every env's asset / executed-action memories on the device so that
        # env_method("save_asset_memory") / ("save_action_memory") return one DataFrame per env
        self._recorder = MemoryRecorder(engine, dates, tickers) if record_memory else None
        # numpy mode: True returns fresh arrays like DummyVecEnv; False returns views of the pinned host buffers
        # (valid until the next step) and saves one host memcpy of the observation batch
        self.copy_outputs = copy_outputs
        self.num_envs = self.env_num = engine.n_envs
        self.tensor_mode = tensor_mode
        D = engine.stock_dim
        if obs_shape is None:
            obs_shape = getattr(engine, "obs_shape", None) or (getattr(engine, "state_space", None) or engine.state_dim,)
        self.obs_shape = tuple(obs_shape)
        self.action_space = Box(low=action_low, high=action_high, shape=(D,), dtype=np.float32)
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=self.obs_shape, dtype=np.float32)
        # ElegantRL's Arguments(env=...) reads these
        self.env_name = getattr(engine, "env_name", type(engine).__name__)
        self.state_dim = int(np.prod(self.obs_shape))
        self.action_dim = D
        self.max_step = getattr(engine, "max_step", getattr(engine, "n_days", 0) - 1)
        self.if_discrete = False
        self.target_return = getattr(engine, "target_return", 10.0)
        self._actions = None
        self._last_obs = None
        if _Base is not object:
            try:
                super().__init__(self.num_envs, self.observation_space, self.action_space)
            except Exception:
                pass

    # ---- SB3 VecEnv protocol -----------------------------------------------------------------
    def reset(self):
        obs = self.engine.reset()
        obs = obs.reshape(self.num_envs, *self.obs_shape)
        self._last_obs = obs
        if self._recorder is not None:
            self._recorder.clear()
        return obs if self.tensor_mode else obs.cpu().numpy()

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        eng = self.engine
        if self.tensor_mode:
            obs, reward, done, flags = eng.step(self._actions, auto_reset=True)
            if self._recorder is not None:
                self._recorder.after_step(flags)
            return obs.reshape(self.num_envs, *self.obs_shape), reward, done, {"flags": flags}
        import torch

        # numpy in / numpy out like DummyVecEnv, through persistent PINNED host buffers: one async upload, one
        # batch of async downloads and a single stream synchronisation per step (two when an env finished)
        a = self._actions
        if isinstance(a, torch.Tensor):
            d_act = a
        else:
            a = np.asarray(a)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            h = self._pinned("act_" + a.dtype.name, (self.num_envs, self.action_dim), torch.from_numpy(a[:0].copy()).dtype)
            np.copyto(h.numpy(), a.reshape(self.num_envs, self.action_dim))
            d_act = h.to(eng.device, non_blocking=True)
        obs, reward, done, flags = eng.step(d_act, auto_reset=False)
        obs_h = self._pinned("obs", (self.num_envs, *self.obs_shape), torch.float32)
        rew_h = self._pinned("rew", (self.num_envs,), reward.dtype)
        flg_h = self._pinned("flags", (self.num_envs,), torch.uint8)
        flg_h.copy_(flags, non_blocking=True)
        rew_h.copy_(reward, non_blocking=True)
        obs_h.copy_(obs.reshape(obs_h.shape), non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        self._flags = flg_h.numpy().copy()
        dones = (self._flags & _cabi.FLAG_DONE).astype(bool)
        rewards = rew_h.numpy().astype(np.float32)
        terminal = {}
        if dones.any():
            idx = np.nonzero(dones)[0]
            term = obs_h.numpy()[idx].copy()
            terminal = {int(i): term[k] for k, i in enumerate(idx)}
            obs = eng.reset(mask=done)  # DummyVecEnv.step_wait: obs = env.reset() for the done envs
            obs_h.copy_(obs.reshape(obs_h.shape), non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
        if self._recorder is not None:
            self._recorder.after_step(flags)
        out = obs_h.numpy()
        return (out.copy() if self.copy_outputs else out), rewards, dones, _LazyInfos(self.num_envs, terminal)

    def _pinned(self, name, shape, dtype):
        import torch

        bufs = self.__dict__.setdefault("_pinned_bufs", {})
        b = bufs.get(name)
        if b is None or tuple(b.shape) != tuple(shape) or b.dtype != dtype:
            b = bufs[name] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
        return b

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed] * self.num_envs

    def render(self, mode="human"):
        return self.engine.observe().reshape(self.num_envs, *self.obs_shape).cpu().numpy()

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            return [indices]
        return [int(i) for i in indices]

    def _per_env(self, value, idx):
        """One value per env: per-env device arrays ([N] / [N, ...] or stock-major [D, N]) are split by env;
        anything else (constructor keywords, tables) is the one setting all N envs share."""
        torch = self.engine._torch
        N = self.num_envs
        if isinstance(value, torch.Tensor) and value.dim() >= 1:
            if value.shape[0] == N:
                host = value.detach().cpu().numpy()
                return [host[i] for i in idx]
            if value.dim() == 2 and value.shape[1] == N:
                host = value.detach().t().cpu().numpy()
                return [host[i] for i in idx]
        return [value for _ in idx]

    def get_attr(self, attr_name, indices=None):
        """SB3 contract: a list with ONE entry per selected env (``VecEnv.get_attr``).  Per-env state arrays of
        the engine (``cash``, ``day``, ``hold`` ...) are returned env by env."""
        return self._per_env(getattr(self.engine, attr_name), self._indices(indices))

    def set_attr(self, attr_name, value, indices=None):
        if indices is not None and len(self._indices(indices)) != self.num_envs:
            raise NotImplementedError("the batched engine shares its settings across envs: set_attr needs indices=None")
        setattr(self.engine, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        """SB3 contract: call the method for every selected env and return the list of results.  The episode
        memories (``save_asset_memory`` / ``save_action_memory``) are kept per env by :class:`MemoryRecorder`
        (``record_memory=True``); other names are engine-level calls whose tensor results are split by env."""
        idx = self._indices(indices)
        rec = self.__dict__.get("_recorder")
        if method_name in ("save_asset_memory", "save_action_memory"):
            if rec is None:
                raise AttributeError(f"{method_name}: construct the VecEnv with record_memory=True (or use the gym "
                                     "class's get_sb_env(), which keeps the reference's memories)")
            return [getattr(rec, method_name)(i) for i in idx]
        r = getattr(self.engine, method_name)(*args, **kwargs)
        return self._per_env(r, idx)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return []


class MemoryRecorder:
    """Per-env ``asset_memory`` / ``actions_memory`` of a batched StockTradingEnv VecEnv (what the reference
    keeps in Python lists, env_stocktrading.py:85-100, :330-352), held as device tensors: one [N] f64 row of
    total assets and one [N, D] i32 row of executed shares per step since the last ``reset()``."""

    def __init__(self, engine, dates=None, tickers=None):
        if getattr(engine, "asset", None) is None:
            raise ValueError("record_memory needs an engine built with track_asset=True")
        self.engine, self.dates, self.tickers = engine, dates, tickers
        self.clear()

    def clear(self):
        """Called after ``reset()`` of all envs: every env's memories restart from its reset state."""
        e = self.engine
        torch = e._torch
        self.base_asset, self.base_day = e.total_asset(), e.day.clone()
        self.start = torch.zeros(e.n_envs, dtype=torch.long, device=e.device)
        self.assets, self.days, self.executed = [], [], []
        self._hold = e.hold.clone()

    def after_step(self, flags):
        """Called once per VecEnv step, after the finished envs were reset (by the kernel or by a masked reset)."""
        e = self.engine
        torch = e._torch
        done = (flags & _cabi.FLAG_DONE) != 0
        self.executed.append((e.hold - self._hold).t().contiguous())  # signed executed shares (:324, :330)
        self.assets.append(e.asset.clone())
        self.days.append(e.day.clone())
        self._hold = e.hold.clone()
        # the terminal call appends nothing (:221-301) and the reset that follows restarts the lists (:364-390)
        self.base_asset = torch.where(done, e.total_asset(), self.base_asset)
        self.base_day = torch.where(done, e.day, self.base_day)
        self.start = torch.where(done, torch.full_like(self.start, len(self.assets)), self.start)

    def _rows(self, i):
        return range(int(self.start[i].item()), len(self.assets))

    def _date(self, d):
        return self.dates[d] if self.dates is not None else d

    def save_asset_memory(self, i):
        import pandas as pd

        rows = self._rows(i)
        days = [int(self.base_day[i].item())] + [int(self.days[k][i].item()) for k in rows]
        vals = [float(self.base_asset[i].item())] + [float(self.assets[k][i].item()) for k in rows]
        return pd.DataFrame({"date": [self._date(d) for d in days], "account_value": vals})

    def save_action_memory(self, i):
        import pandas as pd

        rows = self._rows(i)
        days = [int(self.base_day[i].item())] + [int(self.days[k][i].item()) for k in rows]
        acts = [self.executed[k][i].cpu().numpy().astype(np.int64) for k in rows]
        df = pd.DataFrame(acts, columns=self.tickers)
        df.index = pd.Index([self._date(d) for d in days[:-1]], name="date")
        return df


class GymVecEnv(_Base):
    """SB3's ``DummyVecEnv`` protocol over gym-style env OBJECTS (``DummyVecEnv([lambda: env])``): the envs
    are stepped one after the other, observations land in a float32 buffer, a finished env is reset at once
    and its last observation goes to ``infos[i]["terminal_observation"]``; ``env_method`` / ``get_attr`` /
    ``render`` reach the env objects themselves, which is what ``DRLAgent.DRL_prediction`` and the ensemble
    agent rely on (/root/reference/finrl/agents/stablebaselines3/models.py:110-130, :289-325; the reference's
    ``get_sb_env``, env_stocktrading.py:549-552).  Used when stable-baselines3 is not installed; with SB3
    present the gym classes hand SB3's own ``DummyVecEnv`` the same objects."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        if not self.envs:
            raise ValueError("GymVecEnv needs at least one env")
        e0 = self.envs[0]
        self.num_envs = len(self.envs)
        self.observation_space, self.action_space = e0.observation_space, e0.action_space
        self.metadata = getattr(e0, "metadata", {})
        shape = tuple(self.observation_space.shape)
        self.buf_obs = np.zeros((self.num_envs, *shape), dtype=self.observation_space.dtype)
        self.buf_dones = np.zeros(self.num_envs, dtype=bool)
        self.buf_rews = np.zeros(self.num_envs, dtype=np.float32)
        self.buf_infos = [{} for _ in self.envs]
        self.actions = None
        if _Base is not object:
            try:
                super().__init__(self.num_envs, self.observation_space, self.action_space)
            except Exception:
                pass

    def reset(self):
        for i, e in enumerate(self.envs):
            self.buf_obs[i] = e.reset()
        return self.buf_obs.copy()

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        for i, e in enumerate(self.envs):
            obs, self.buf_rews[i], self.buf_dones[i], info = e.step(self.actions[i])
            if self.buf_dones[i]:
                info = dict(info)
                info["terminal_observation"] = obs
                obs = e.reset()
            self.buf_infos[i] = info
            self.buf_obs[i] = obs
        return self.buf_obs.copy(), self.buf_rews.copy(), self.buf_dones.copy(), [dict(i) for i in self.buf_infos]

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        for e in self.envs:
            if hasattr(e, "close"):
                e.close()

    def seed(self, seed=None):
        out = []
        for i, e in enumerate(self.envs):
            s = None if seed is None else seed + i
            fn = getattr(e, "seed", None) or getattr(e, "_seed", None)
            out.append(fn(s) if fn is not None else None)
        return out

    def render(self, mode="human"):
        if self.num_envs == 1:
            return self.envs[0].render(mode=mode)
        return [e.render(mode=mode) for e in self.envs]

    def _targets(self, indices):
        if indices is None:
            return self.envs
        if isinstance(indices, int):
            return [self.envs[indices]]
        return [self.envs[i] for i in indices]

    def get_attr(self, attr_name, indices=None):
        return [getattr(e, attr_name) for e in self._targets(indices)]

    def set_attr(self, attr_name, value, indices=None):
        for e in self._targets(indices):
            setattr(e, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        return [getattr(e, method_name)(*args, **kwargs) for e in self._targets(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._targets(indices)]

    def get_images(self):
        return []


def dummy_vec_env(env_fns):
    """``DummyVecEnv(env_fns)``: stable-baselines3's class when it is installed (what the reference imports),
    :class:`GymVecEnv` otherwise."""
    try:
        from stable_baselines3.common.vec_env import DummyVecEnv  # type: ignore
    except Exception:
        return GymVecEnv(env_fns)
    return DummyVecEnv(env_fns)


class _DeepCopyVecMixin:
    """``get_sb_env`` of the cash-penalty family: ``DummyVecEnv([lambda: deepcopy(self)])`` + first observation
    (env_stocktrading_cashpenalty.py:374-380).  ``deepcopy`` of a drop-in gives the copy its own 1-env engine
    on the shared (read-only) tables with the per-env device state cloned, and copies the Python memories."""

    def __deepcopy__(self, memo):
        import copy

        new = type(self).__new__(type(self))
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "engine":
                continue
            new.__dict__[k] = v if k in ("df", "_device") else copy.deepcopy(v, memo)
        new.engine = self._make_engine(1)
        self.engine.clone_state_into(new.engine)
        return new

    def get_sb_env(self):
        import copy

        e = dummy_vec_env([lambda: copy.deepcopy(self)])
        obs = e.reset()
        return e, obs
