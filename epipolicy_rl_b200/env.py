"""Drop-in twins of the reference's gym environment (epipolicy_environment.py:9-71).

`EpiEnv`        — same constructor shape, attributes and reset/step/render/close surface as the
                  reference class, one environment, float64 numpy observations (so SB3's
                  DummyVecEnv/Monitor, baselines.py and plots.py can use it unchanged).
`BatchedEpiEnv` — N environments stepped by one fused launch; torch CUDA tensors in and out.
`EpiVecEnv`     — the batched environment behind stable-baselines3's VecEnv surface (numpy, auto-reset).

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


try:  # pragma: no cover - stable_baselines3 is not in this image; with it the adapter below is a real VecEnv
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:  # noqa: BLE001
    _VecEnvBase = object


class EpiVecEnv(_VecEnvBase):
    """The batched environment behind stable-baselines3's VecEnv surface, so that the reference's
    runner.py:62-155 can train on N GPU-resident envs instead of its DummyVecEnv of one: numpy in and out,
    `step_async` / `step_wait`, auto-reset with the terminal observation in `infos[i]["terminal_observation"]`,
    `num_envs`, `get_attr / set_attr / env_method / env_is_wrapped / seed` as VecEnv defines them. One host round trip
    per gym step of ALL envs (the engine's pinned host buffers); the in-repo learner (learner.py) avoids even that.

    `batched`: a BatchedEpiEnv, or any object with `.epi` offering step_host(actions) -> (obs, reward, done) numpy
    views, reset(mask=None), obs_dim / action_dim, and `.action_space / .observation_space`."""

    def __init__(self, batched):
        self.batched, self.epi = batched, batched.epi
        self.num_envs = int(getattr(batched, "n_envs", getattr(self.epi, "n_envs", 1)))
        self.observation_space, self.action_space = batched.observation_space, batched.action_space
        if _VecEnvBase is not object:  # pragma: no cover
            super().__init__(self.num_envs, self.observation_space, self.action_space)
        self._actions = None
        self.render_mode = None

    def _obs_now(self):
        o = self.epi.obs
        return np.asarray(o.double().cpu().numpy() if hasattr(o, "cpu") else o, np.float64)

    def reset(self):
        self.epi.reset()
        return self._obs_now().copy()

    def step_async(self, actions):
        a = np.asarray(actions, np.float32).reshape(self.num_envs, -1)
        if a.shape[1] < self.epi.action_dim:
            raise ValueError(f"actions have {a.shape[1]} entries per env, the scenario has {self.epi.action_dim} free control parameters")
        self._actions = np.ascontiguousarray(a[:, :self.epi.action_dim])  # the reference ignores the surplus entries

    def step_wait(self):
        obs, rew, done = self.epi.step_host(self._actions, PERIOD)
        obs, rew, done = np.array(obs, np.float64), np.array(rew, np.float64), np.array(done).astype(bool)
        infos = [{} for _ in range(self.num_envs)]
        if done.any():
            for i in np.nonzero(done)[0]:
                infos[i]["terminal_observation"] = obs[i].copy()
            self.epi.reset(_mask_for(self.epi, done))
            fresh = self._obs_now()
            obs[done] = fresh[done]
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if hasattr(self.batched, "close"):
            self.batched.close()

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.batched, attr_name)] * len(self._indices(indices))

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.batched, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return [getattr(self.batched, method_name)(*method_args, **method_kwargs)] * len(self._indices(indices))

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * len(self._indices(indices))

    def seed(self, seed=None):
        return [None] * self.num_envs  # the simulator is deterministic

    def render(self, mode="human"):
        print("Rendering")

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        return [indices] if isinstance(indices, int) else list(indices)


def _mask_for(epi, done):
    """reset mask in the form the engine takes (a torch tensor for the GPU engine, numpy for host-side doubles)"""
    if hasattr(epi, "device"):
        return torch.from_numpy(done.astype(np.uint8))
    return done.astype(np.uint8)


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
