"""Golden fixtures for the part of the intervention algebra the shipped schedules never reach:
mobility multipliers with a NON-EMPTY locale selection (`sim.apply({'locale-from':..,'locale-to':..}, m)`,
core/executor.py:121-134 -> CooMatrix.pairwise_multiply + the row renormalisation of matrix/coo.py:173-188).

The reference CODE is unmodified; the scenario JSON is COVID_C.json with its `Border_Closures` schedule pointed at one
real locale (so that '~' + locales is not empty) and given two phases (degree 0.3 -> 0.6 on day 59, so the mobility
gap is non-zero on a mid-run day as well as on day 0).

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_border_golden.py
-> tests/golden/desc_COVID_C_border.npz, tests/golden/traj_COVID_C_border.npz
"""
import copy
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402  (patches solve_ivp for the nfev trace, imports the reference)
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200.export import export_static, export_initial_parameter  # noqa: E402

T_DAYS = 90


def border_session():
    s = json.load(open(os.path.join(gg.REF, "jsons", "COVID_C.json")))
    for sch in s["schedules"]:
        if sch["name"] != "Border_Closures":
            continue
        d0 = sch["detail"][0]
        d0["locales"] = "UnitedProvinces.Beaches"
        d0["end_date"] = "2021-02-28"
        d1 = copy.deepcopy(d0)
        d1["id"] = 2
        d1["start_date"], d1["end_date"] = "2021-03-01", "2021-12-31"
        d1["control_params"][0]["value"] = "0.6"
        sch["detail"].append(d1)
    return s


def main():
    t0 = time.time()
    name = "COVID_C_border"
    env = EpiEnv(border_session())
    epi = env.epi
    d = export_static(epi)
    d.meta["scenario"] = name
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    out = {f"init_{k[5:]}": v for k, v in export_initial_parameter(epi).items()}
    for k, v in gg.run_schedule_plan(epi, T_DAYS).items():
        out[f"sched_{k}"] = v
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **out)
    kinds = list(np.asarray(d.op_kind))
    print(f"{name}: done in {time.time() - t0:.1f}s  ops={d.NOP} MOB_MUL ops={kinds.count(4)}", flush=True)


if __name__ == "__main__":
    main()
