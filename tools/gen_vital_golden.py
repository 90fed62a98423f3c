"""Golden fixture for NON-CONSERVING model terms: an equation term with no counterpart in another compartment becomes a
`gain_map` / `lost_map` entry of its edge (obj/edge.py:76-111) and feeds `cumulative_gain` / `cumulative_lost`
(core_optimize.py:171-195), which are part of the flat ODE state and of the RK45 error norm. Every shipped model
conserves mass, so these paths never run on the shipped scenarios.

The reference CODE is unmodified; the scenario JSON is SIRV_B.json with births into S (`+ mu * S`) and deaths out of R
(`- mu * R`) added to its model.

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_vital_golden.py
-> tests/golden/desc_SIRV_B_vital.npz, tests/golden/traj_SIRV_B_vital.npz
"""
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200.export import export_static, export_initial_parameter  # noqa: E402

N_WEEKS, T_SCHED = 20, 30


def vital_session():
    s = json.load(open(os.path.join(gg.REF, "jsons", "SIRV_B.json")))
    m = s["model"]
    m["parameters"].append({"id": len(m["parameters"]) + 1, "name": "mu", "desc": "birth / death rate", "default_value": "0.004", "tags": []})
    for c in m["compartments"]:
        if c["name"] == "S":
            c["equation"] = c["equation"] + " + mu * S"
        if c["name"] == "R":
            c["equation"] = c["equation"] + " - (mu * R)"
    return s


def main():
    t0 = time.time()
    name = "SIRV_B_vital"
    env = EpiEnv(vital_session())
    epi = env.epi
    d = export_static(epi)
    d.meta["scenario"] = name
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    out = {f"init_{k[5:]}": v for k, v in export_initial_parameter(epi).items()}
    A = env.action_space.shape[0]
    acts = np.random.default_rng(13).uniform(0, 1, (N_WEEKS, A)).astype(np.float32)
    for k, v in gg.run_env_plan(env, acts).items():
        out[f"rnd_{k}"] = v
    # the flat observable of the last state: cumulative_gain / cumulative_lost are in it (obj/observable.py:24-25)
    out["rnd_flat_last"] = np.asarray(epi.history.states[-1].obs.flatten())
    for k, v in gg.run_schedule_plan(epi, T_SCHED).items():
        out[f"sched_{k}"] = v
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **out)
    print(f"{name}: done in {time.time() - t0:.1f}s  n_flat={d.n_flat} sc_kind={sorted(set(int(k) for k in d.arrays['sc_kind']))}", flush=True)


if __name__ == "__main__":
    main()
