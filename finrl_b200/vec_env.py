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
                 copy_outputs: bool = True):
        self.engine = engine
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
        return obs if self.tensor_mode else obs.cpu().numpy()

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        eng = self.engine
        if self.tensor_mode:
            obs, reward, done, flags = eng.step(self._actions, auto_reset=True)
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

    def get_attr(self, attr_name, indices=None):
        v = getattr(self.engine, attr_name)
        n = self.num_envs if indices is None else len(list(indices))
        return [v] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.engine, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        r = getattr(self.engine, method_name)(*args, **kwargs)
        n = self.num_envs if indices is None else len(list(indices))
        return [r] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(list(indices))
        return [False] * n

    def get_images(self):
        return []
