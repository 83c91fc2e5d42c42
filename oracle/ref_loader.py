"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference env files from /root/reference.

Used by ``tests/golden/make_golden.py`` (fixture generation, in the build container) and by
tests that are skipped when /root/reference is absent.  Nothing in ``finrl_b200/`` imports this.

The reference's env modules import ``gym``, ``matplotlib`` and ``stable_baselines3``, none of
which is installed and none of which touches the step arithmetic; ``finrl/__init__.py`` pulls
alpaca/elegantrl/ray.  We therefore inject inert stub modules and load the four env files
directly with ``importlib.util.spec_from_file_location``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("FINRL_REFERENCE_ROOT", "/root/reference")

ENV_FILES = {
    "env_stocktrading": "finrl/meta/env_stock_trading/env_stocktrading.py",
    "env_stocktrading_np": "finrl/meta/env_stock_trading/env_stocktrading_np.py",
    "env_stocktrading_cashpenalty": "finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py",
    "env_portfolio": "finrl/meta/env_portfolio_allocation/env_portfolio.py",
    "env_nas100_wrds": "finrl/meta/env_stock_trading/env_nas100_wrds.py",
    "env_multiple_crypto": "finrl/meta/env_cryptocurrency_trading/env_multiple_crypto.py",
    "env_stocktrading_stoploss": "finrl/meta/env_stock_trading/env_stocktrading_stoploss.py",
    "preprocessors": "finrl/meta/preprocessor/preprocessors.py",
}


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_ROOT, p)) for p in ENV_FILES.values())


class _Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        import numpy as np

        self.low, self.high, self.shape = low, high, tuple(shape) if shape is not None else None
        self.dtype = np.dtype(dtype if dtype is not None else np.float32)


class _StubDummyVecEnv:
    """Minimal restatement of SB3's DummyVecEnv protocol (sequential loop, auto-reset,
    float32 observation buffer); enough for ``get_sb_env`` and the CPU timing baseline."""

    def __init__(self, env_fns):
        import numpy as np

        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        shape = self.envs[0].observation_space.shape
        self.buf_obs = np.zeros((self.num_envs,) + tuple(shape), dtype=np.float32)

    def reset(self):
        for i, e in enumerate(self.envs):
            self.buf_obs[i] = e.reset()
        return self.buf_obs.copy()

    def step(self, actions):
        import numpy as np

        rews = np.zeros(self.num_envs, dtype=np.float32)
        dones = np.zeros(self.num_envs, dtype=bool)
        infos = []
        for i, e in enumerate(self.envs):
            obs, rews[i], dones[i], info = e.step(actions[i])
            if dones[i]:
                info = dict(info)
                info["terminal_observation"] = obs
                obs = e.reset()
            self.buf_obs[i] = obs
            infos.append(info)
        return self.buf_obs.copy(), rews, dones, infos

    def env_method(self, method_name, *a, indices=None, **k):
        return [getattr(e, method_name)(*a, **k) for e in self.envs]

    def render(self, mode="human"):
        return self.envs[0].render(mode=mode)


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # noqa: D401
            pass

        gym.Env = Env
        spaces = types.ModuleType("gym.spaces")
        spaces.Box = _Box
        gym.spaces = spaces
        utils = types.ModuleType("gym.utils")
        seeding = types.ModuleType("gym.utils.seeding")

        def np_random(seed=None):
            import numpy as np

            return np.random.RandomState(seed), seed

        seeding.np_random = np_random
        utils.seeding = seeding
        gym.utils = utils
        logger = types.ModuleType("gym.logger")
        logger.set_level = lambda *_a, **_k: None
        gym.logger = logger
        sys.modules.update(
            {"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding, "gym.logger": logger}
        )
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *_a, **_k: None
        plt = types.ModuleType("matplotlib.pyplot")
        for name in ("plot", "savefig", "close", "figure", "show"):
            setattr(plt, name, lambda *_a, **_k: None)
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    if "stable_baselines3" not in sys.modules:
        sb3 = types.ModuleType("stable_baselines3")
        common = types.ModuleType("stable_baselines3.common")
        vec = types.ModuleType("stable_baselines3.common.vec_env")
        vec.DummyVecEnv = _StubDummyVecEnv
        vec.SubprocVecEnv = _StubDummyVecEnv
        lg = types.ModuleType("stable_baselines3.common.logger")
        lg.record = lambda *_a, **_k: None
        common.vec_env = vec
        common.logger = lg
        sb3.common = common
        sys.modules.update(
            {
                "stable_baselines3": sb3,
                "stable_baselines3.common": common,
                "stable_baselines3.common.vec_env": vec,
                "stable_baselines3.common.logger": lg,
            }
        )


def _install_finrl_stubs():
    """env_multiple_crypto.py imports the agent adapters and DataProcessor at module level (unused by
    the env class); give it inert namespaces instead of importing ray / elegantrl / alpaca."""
    if "finrl" in sys.modules:
        return
    names = ["finrl", "finrl.agents", "finrl.agents.elegantrl", "finrl.agents.elegantrl.models",
             "finrl.agents.stablebaselines3", "finrl.agents.stablebaselines3.models", "finrl.meta",
             "finrl.meta.data_processor", "finrl.config", "finrl.meta.preprocessor",
             "finrl.meta.preprocessor.yahoodownloader", "stockstats"]
    for n in names:
        sys.modules[n] = types.ModuleType(n)
    # preprocessors.py: `from stockstats import StockDataFrame`, `from finrl import config`, YahooDownloader —
    # module-level imports that calculate_turbulence / data_split never touch
    sys.modules["stockstats"].StockDataFrame = object
    sys.modules["finrl"].config = sys.modules["finrl.config"]
    sys.modules["finrl.config"].INDICATORS = []
    sys.modules["finrl.meta.preprocessor.yahoodownloader"].YahooDownloader = object
    sys.modules["finrl.agents.elegantrl.models"].DRLAgent = object
    sys.modules["finrl.agents.stablebaselines3.models"].DRLAgent = object
    sys.modules["finrl.meta.data_processor"].DataProcessor = object


_cache: dict = {}


def load(name: str):
    """Return the reference module ``name`` (a key of ENV_FILES), loaded unmodified."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    _install_stubs()
    if name in ("env_multiple_crypto", "preprocessors"):
        _install_finrl_stubs()
    path = os.path.join(REF_ROOT, ENV_FILES[name])
    spec = importlib.util.spec_from_file_location("_finrl_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod
