"""TEST INFRASTRUCTURE — the reference's own SB3 adapter loops, restated so that they can drive either the
unmodified reference env classes (tests/golden/make_golden.py, build container) or the finrl_b200 drop-ins
(tests/test_adapter_flows_gpu.py).  The agent module itself cannot be imported (it needs stable-baselines3),
so the two loops that touch the envs are written out here, statement for statement:

* ``drl_prediction``  = ``DRLAgent.DRL_prediction``
  (/root/reference/finrl/agents/stablebaselines3/models.py:110-130)
* ``ensemble_trade_window`` = the env part of ``DRLEnsembleAgent.DRL_prediction`` (models.py:278-325): a
  ``DummyVecEnv([lambda: StockTradingEnv(..., initial=..., previous_state=last_state, mode="trade")])`` stepped
  over one window, ``last_state = trade_env.render()`` taken on the last-but-one step.
"""
from __future__ import annotations

import numpy as np


class ReplayModel:
    """Stands in for an SB3 model: ``predict`` replays a fixed action table ([steps, D]) as [1, D] batches."""

    def __init__(self, actions):
        self.actions, self.calls = np.asarray(actions), 0

    def predict(self, obs, deterministic=True):
        assert obs.shape[0] == 1 and obs.dtype == np.float32
        a = self.actions[self.calls % len(self.actions)]
        self.calls += 1
        return a[None, :].copy(), None


def drl_prediction(model, environment, deterministic=True):
    test_env, test_obs = environment.get_sb_env()
    account_memory = []
    actions_memory = []
    test_env.reset()
    for i in range(len(environment.df.index.unique())):
        action, _states = model.predict(test_obs, deterministic=deterministic)
        test_obs, rewards, dones, info = test_env.step(action)
        if i == (len(environment.df.index.unique()) - 2):
            account_memory = test_env.env_method(method_name="save_asset_memory")
            actions_memory = test_env.env_method(method_name="save_action_memory")
        if dones[0]:
            break
    return account_memory[0], actions_memory[0]


def ensemble_trade_window(make_vec, env_cls, trade_data, model, last_state, initial, env_kwargs, name, iter_num, trace=None):
    """``make_vec(list_of_callables)`` is the DummyVecEnv constructor in use.  ``trace`` (a dict) collects what
    the loop saw, step by step, for the goldens."""
    trade_env = make_vec([
        lambda: env_cls(df=trade_data, initial=initial, previous_state=last_state, model_name=name, mode="trade",
                        iteration=iter_num, **env_kwargs)
    ])
    trade_obs = trade_env.reset()
    if trace is not None:
        trace.update(obs0=trade_obs[0].copy(), obs=[], rewards=[], dones=[])
    for i in range(len(trade_data.index.unique())):
        action, _states = model.predict(trade_obs)
        trade_obs, rewards, dones, info = trade_env.step(action)
        if trace is not None:
            trace["obs"].append(trade_obs[0].copy())
            trace["rewards"].append(rewards[0])
            trace["dones"].append(dones[0])
        if i == (len(trade_data.index.unique()) - 2):
            last_state = trade_env.render()
    return last_state
