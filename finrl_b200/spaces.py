"""gym / gymnasium / stable-baselines3 are optional: none of them is part of the step path.

``Box`` resolves to gymnasium's or gym's class when one is importable (so SB3 accepts the spaces) and
to a minimal stand-in otherwise; ``VecEnvBase`` likewise resolves to SB3's ``VecEnv`` when present.
"""
from __future__ import annotations

import numpy as np


def _resolve_box():
    for mod in ("gymnasium", "gym"):
        try:
            m = __import__(mod)
            return m.spaces.Box, mod
        except Exception:
            continue
    return None, None


_Box, BOX_SOURCE = _resolve_box()


class _FallbackBox:
    """Just enough of gym.spaces.Box for shape/dtype/bounds queries and sampling."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def Box(low, high, shape=None, dtype=np.float32):
    if _Box is not None:
        return _Box(low=low, high=high, shape=shape, dtype=dtype)
    return _FallbackBox(low, high, shape, dtype)


def vec_env_base():
    try:
        from stable_baselines3.common.vec_env import VecEnv  # type: ignore

        return VecEnv
    except Exception:
        return object


def gym_env_base():
    """``gym.Env`` (or gymnasium's) when importable, so the drop-in classes are real ``Env`` subclasses for
    SB3's wrappers and type checks; ``object`` otherwise (none of them is installed in the build image)."""
    for mod in ("gym", "gymnasium"):
        try:
            return __import__(mod).Env
        except Exception:
            continue
    return object
