"""EpiVecEnv (epipolicy_rl_b200/env.py): the SB3 VecEnv surface over the batched engine, on the emulated kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from emu_engine import EmuEngine  # noqa: E402
from helpers import load  # noqa: E402

from epipolicy_rl_b200.env import Box, EpiVecEnv  # noqa: E402


class _HostBatched:
    """what EpiVecEnv needs of a BatchedEpiEnv, on the host emulation"""

    def __init__(self, d, n):
        e = EmuEngine(d, n, "fp64")
        e.n_envs, e.obs_dim, e.action_dim = n, e.info.obs_dim, e.info.action_dim
        e.step_host = lambda a, days: e.step(a, days)
        self.epi, self.n_envs = e, n
        self.action_space = Box(0, 1, (d.A,), np.float32)
        self.observation_space = Box(0, float(np.sum(d.X0)), (e.info.obs_dim,), np.float64)

    def close(self):
        self.epi.close()


def test_vecenv_surface_and_auto_reset():
    d, _ = load("SIR_A")
    n = 3
    venv = EpiVecEnv(_HostBatched(d, n))
    assert venv.num_envs == n and venv.action_space.shape == (d.A,)
    obs0 = venv.reset()
    assert obs0.shape == (n, d.C * d.L * d.G) and obs0.dtype == np.float64
    # start env 2 one week before the horizon: it finishes on the first step and is restarted in place
    day = venv.epi.get_state("day")
    rng = np.random.default_rng(0)
    a = rng.uniform(0, 1, (n, d.A + 2)).astype(np.float32)  # surplus action entries are ignored, as in the reference
    steps_to_end = (d.horizon + 6) // 7
    seen_done = False
    for k in range(steps_to_end):
        venv.step_async(a)
        obs, rew, done, infos = venv.step_wait()
        assert obs.shape == obs0.shape and rew.shape == (n,) and done.dtype == bool and len(infos) == n
        if done.any():
            seen_done = True
            assert done.all()  # identical horizons
            for i in range(n):
                assert "terminal_observation" in infos[i]
                assert not np.array_equal(infos[i]["terminal_observation"], obs[i])
            assert np.array_equal(obs, obs0)           # the returned observation is the fresh episode's
            assert (venv.epi.get_state("day") == 0).all()
    assert seen_done and (day == 0).all()
    assert venv.env_is_wrapped(object) == [False] * n and venv.get_attr("n_envs") == [n] * n
    venv.close()
