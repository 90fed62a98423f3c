"""Golden fixture for the renormalisation branch of facility_interaction (get_normalized_facility_interaction,
matrix/dense.py:21-32: a row is rescaled only when its sum is > 1 or < 0), which the shipped schedules never reach
because their multipliers are <= 1. The reference CODE is unmodified; the scenario JSON is COVID_C.json with the
`School closure` and `Workplace closure` schedules given NEGATIVE percentages (multipliers 1.6 and 1.3: more contact
than the default), switched to other values on day 20.

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_finorm_golden.py
-> tests/golden/desc_COVID_C_finorm.npz, tests/golden/traj_COVID_C_finorm.npz
"""
import copy
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200.export import export_static, export_initial_parameter  # noqa: E402

T_SCHED = 45


def finorm_session():
    s = json.load(open(os.path.join(gg.REF, "jsons", "COVID_C.json")))
    for sch in s["schedules"]:
        if sch["name"] in ("School closure", "Workplace closure"):
            first = sch["name"] == "School closure"
            d0 = sch["detail"][0]
            d0["control_params"][0]["value"] = "-0.6" if first else "-0.3"
            d0["end_date"] = "2021-01-20"
            d1 = copy.deepcopy(d0)
            d1["id"] = 2
            d1["start_date"], d1["end_date"] = "2021-01-21", "2021-12-31"
            d1["control_params"][0]["value"] = "-0.2" if first else "-0.1"
            sch["detail"].append(d1)
    return s


def main():
    t0 = time.time()
    name = "COVID_C_finorm"
    env = EpiEnv(finorm_session())
    epi = env.epi
    d = export_static(epi)
    d.meta["scenario"] = name
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    out = {f"init_{k[5:]}": v for k, v in export_initial_parameter(epi).items()}
    for k, v in gg.run_schedule_plan(epi, T_SCHED).items():
        out[f"sched_{k}"] = v
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **out)
    fi0 = np.asarray(d.arrays["fi0"]).reshape(d.L, d.F, d.G, d.G)
    print(f"{name}: done in {time.time() - t0:.1f}s; default fi row sums max {fi0.sum(-1).max():.3f}; "
          f"day-0 fi row sums max {out['sched_p_fi'][0].sum(-1).max():.3f}", flush=True)


if __name__ == "__main__":
    main()
