"""Drop-in twins of the reference's gym environment (epipolicy_environment.py:9-71).

`EpiEnv`        — same constructor shape, attributes and reset/step/render/close surface as the
                  reference class, one environment, float64 numpy observations (so SB3's
                  DummyVecEnv/Monitor, baselines.py and plots.py can use it unchanged).
`BatchedEpiEnv` — N environments stepped by one fused launch; torch CUDA tensors in and out.

Construction takes either an `EpiDesc` (no reference needed at run time) or a reference session dict
(then the reference's own parser / make_static build the Static and export.py flattens it).
"""
from __future__ import annotations

import numpy as np
import torch

from .desc import EpiDesc
from .engine import PERIOD, BatchedEpidemic


class Box:
    """What the reference uses of gym.spaces.Box; replaced by the real class when gym is installed."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        return np.random.uniform(self.low, self.high, self.shape).astype(self.dtype)


def _box(low, high, shape, dtype):
    try:  # pragma: no cover - gym is not in this image
        from gym import spaces
        return spaces.Box(low=low, high=high, shape=shape, dtype=dtype)
    except Exception:  # noqa: BLE001
        return Box(low, high, shape, dtype)


def desc_from_session(session) -> EpiDesc:
    """JSON session -> EpiDesc through the reference's own construct_epidemic (kept as is)."""
    from epipolicy.core.epidemic import construct_epidemic  # the reference package must be importable
    from .export import export_static
    return export_static(construct_epidemic(session))


class BatchedEpiEnv:
    def __init__(self, scenario, n_envs: int, device=0, precision: str = "fp32", auto_reset: bool = False):
        self.desc = scenario if isinstance(scenario, EpiDesc) else desc_from_session(scenario)
        self.epi = BatchedEpidemic(self.desc, n_envs, device=device, precision=precision)
        self.n_envs, self.auto_reset = n_envs, auto_reset
        self.act_domain = self.desc.act_domain
        n_act = int(self.desc.meta.get("n_action_interventions", self.desc.A))
        self.action_space = _box(0, 1, (n_act,), np.float32)
        self.observation_space = _box(0, float(self.desc.meta.get("total_population", np.sum(self.desc.X0))),
                                      (self.epi.obs_dim,), np.float64)

    def reset(self, mask=None):
        return self.epi.reset(mask)

    def step(self, actions: torch.Tensor):
        obs, rew, done = self.epi.step(actions, PERIOD)
        if self.auto_reset:
            # terminal observation is returned; the env restarts for the next call (VecEnv convention)
            obs = obs.clone()
            self.epi.reset(done)
        return obs, rew, done, {}

    def close(self):
        self.epi.close()


try:  # pragma: no cover - gym is not in this image; with it the twin is a real gym.Env (SB3's Monitor checks)
    import gym as _gym
    _EnvBase = _gym.Env
except Exception:  # noqa: BLE001
    _EnvBase = object


class EpiEnv(_EnvBase):
    """Single-environment twin of the reference's EpiEnv."""

    metadata = {"render.modes": ["human"]}

    def __init__(self, session, device=0, precision: str = "fp64"):
        self._b = BatchedEpiEnv(session, 1, device=device, precision=precision)
        self.epi = self._b.epi
        self.act_domain = self._b.act_domain
        self.action_space = self._b.action_space
        self.observation_space = self._b.observation_space

    def step(self, action):
        # the reference consumes the first n_free entries of the action and ignores the rest
        # (epipolicy_environment.py:40-47): action_space counts non-cost interventions, which can exceed the number of
        # free control parameters (warmup_scenario: 2 vs 0)
        a = np.asarray(action, np.float32).reshape(-1)
        if a.shape[0] < self.epi.action_dim:
            raise ValueError(f"action has {a.shape[0]} entries, the scenario has {self.epi.action_dim} free control parameters")
        obs, rew, done = self.epi.step_host(a[:self.epi.action_dim].reshape(1, -1), PERIOD)
        return obs[0].astype(np.float64), float(rew[0]), bool(done[0]), dict()

    def reset(self):
        return self.epi.reset()[0].double().cpu().numpy()

    def render(self, mode="human"):
        print("Rendering")

    def close(self):
        pass
