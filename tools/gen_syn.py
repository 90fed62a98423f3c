"""Synthetic scenarios (epipolicy_rl_b200/synth.py) through the UNMODIFIED reference: exports the Static the
reference's parser/make_static build (tests/golden/desc_syn_*.npz) and a short reference trajectory
(tests/golden/traj_syn_*.npz) for parity tests at L*G > 32.

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:. python tools/gen_syn.py syn_s syn_m syn
"""
import json
import os
import sys
import time

import numpy as np

REF = os.environ.get("EPI_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
SCN = os.path.join(os.path.dirname(__file__), "..", "epipolicy_rl_b200", "scenarios")  # product data: the descriptors

import epipolicy.core.epidemic as ref_epidemic  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402

from epipolicy_rl_b200 import synth  # noqa: E402
from epipolicy_rl_b200.export import export_static  # noqa: E402

_trace = []
_orig = ref_epidemic.solve_ivp


def _traced(*a, **k):
    res = _orig(*a, **k)
    _trace.append((res.nfev, len(res.t) - 1, res.status))
    return res


ref_epidemic.solve_ivp = _traced
N_STEPS = {"syn_s": 4, "syn_m": 2, "syn": 1}


def main(names):
    base = json.load(open(os.path.join(REF, "jsons", "COVID_C.json")))
    for name in names:
        t0 = time.time()
        L, G, F = synth.SIZES[name]
        session = synth.make_session(base, L, G, F)
        env = EpiEnv(session)
        d = export_static(env.epi)
        d.meta["scenario"] = name
        d.save(os.path.join(SCN, f"desc_{name}.npz"))
        print(f"{name}: constructed + exported in {time.time() - t0:.1f}s (L={d.L} G={d.G} F={d.F} nnz={d.nnz})", flush=True)
        A = env.action_space.shape[0]
        acts = np.random.default_rng(5).uniform(0.05, 0.6, (N_STEPS[name], A)).astype(np.float32)
        _trace.clear()
        env.reset()
        rewards, obs_w = [], []
        for a in acts:
            t1 = time.time()
            obs, r, done, _ = env.step(a)
            rewards.append(r)
            obs_w.append(obs)
            print(f"  gym step: {time.time() - t1:.1f}s reward {r!r}", flush=True)
        st = env.epi.history.states
        np.savez_compressed(os.path.join(OUT, f"traj_{name}.npz"),
                            actions=acts, rewards=np.array(rewards), obs_weekly=np.stack(obs_w),
                            X=np.stack([np.array(s.obs.current_comp) for s in st]).astype(np.float64),
                            cost=np.stack([np.array(s.obs.cumulative_cost) for s in st]),
                            nfev=np.array([t[0] for t in _trace], np.int32))
        print(f"{name}: done in {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["syn_s", "syn_m"])
