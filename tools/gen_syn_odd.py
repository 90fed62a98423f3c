"""A synthetic scenario with AWKWARD dimensions through the unmodified reference: 13 locales x 5 groups x 3 facilities
(L*G = 65 > 32: env / env-group kernels; G neither a multiple of 4 nor 16, F odd, L prime - the shipped synthetic sizes
are 20x4x8, 100x8x8 and 1000x16x8, so the kernels' vectorised paths for G % 4 == 0, G == 16, F % 2 == 0 are the only ones
they exercise).

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_syn_odd.py
-> tests/golden/desc_syn_odd.npz, traj_syn_odd.npz; desc_syn_cell.npz, traj_syn_cell.npz (5 x 3 x 3: the cell kernel with
   more than four locales and two envs per warp)
"""
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200 import synth  # noqa: E402
from epipolicy_rl_b200.export import export_static  # noqa: E402


def main(name="syn_odd", dims=(13, 5, 3)):
    t0 = time.time()
    L, G, F = dims
    base = json.load(open(os.path.join(gg.REF, "jsons", "COVID_C.json")))
    env = EpiEnv(synth.make_session(base, L, G, F, seed=20260202))
    d = export_static(env.epi)
    d.meta["scenario"] = name
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    A = env.action_space.shape[0]
    acts = np.random.default_rng(6).uniform(0.05, 0.9, (4, A)).astype(np.float32)
    out = gg.run_env_plan(env, acts)
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **{f"rnd_{k}": v for k, v in out.items()})
    print(f"{name}: done in {time.time() - t0:.1f}s (L={d.L} G={d.G} F={d.F} nnz={d.nnz})", flush=True)


if __name__ == "__main__":
    main()
    # the cell kernel's shape space: L*G <= 32 with more than 4 locales (its dense-register mobility covers L <= 4 only;
    # every shipped scenario has L <= 3) and two envs per warp
    main("syn_cell", (5, 3, 3))
    main("syn_lg32", (2, 16, 3))  # L*G == 32: a full warp per env, the largest scenario of the cell kernel
