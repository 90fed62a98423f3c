import os

import numpy as np

from epipolicy_rl_b200.desc import EpiDesc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    """(descriptor, golden trajectories); test-only scenario variants keep their descriptor next to the trajectories"""
    from epipolicy_rl_b200 import scenarios
    local = os.path.join(GOLDEN, f"desc_{name}.npz")
    d = EpiDesc.load(local) if os.path.exists(local) else scenarios.load(name)
    return d, np.load(os.path.join(GOLDEN, f"traj_{name}.npz"))


def x_err(x, ref, desc):
    """SURVEY.md §8(d) parity metric: |x - ref| / (|ref| + 1e-6 * pop[l,g]), max over elements."""
    pop = desc.pop.reshape(1, desc.L, desc.G)
    x = np.asarray(x, np.float64).reshape(-1, desc.C, desc.L, desc.G)
    ref = np.asarray(ref, np.float64).reshape(-1, desc.C, desc.L, desc.G)
    return float(np.max(np.abs(x - ref) / (np.abs(ref) + 1e-6 * pop)))


def cost_err(c, ref):
    c, ref = np.asarray(c, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(c - ref) / (np.abs(ref) + 1e-9 * max(np.max(np.abs(ref)), 1e-300))))


def expand_actions(desc, actions):
    """EpiEnv.step's action expansion (epipolicy_environment.py:40-54) -> (cpv [.., NCP], active [I])."""
    actions = np.asarray(actions, np.float32)
    cpv = np.tile(desc.cp_default.astype(np.float32), actions.shape[:-1] + (1,))
    a = 0
    for i in range(desc.I):
        if desc.itv_is_cost[i]:
            continue
        for k in range(desc.itv_cp_off[i], desc.itv_cp_off[i + 1]):
            if desc.cp_min[k] == desc.cp_max[k]:
                cpv[..., k] = desc.cp_min[k]
            else:
                cpv[..., k] = actions[..., a]
                a += 1
    return cpv
