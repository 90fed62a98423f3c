"""Golden fixture for CROSS-LOCALE moves (nb_move's branches for targets in other locales, core/executor.py:185-200 and
:213-233; applied with per-source clamping at core_optimize.py:196-208), which no shipped intervention uses.

The reference CODE is unmodified; the scenario JSON is COVID_C.json with
  * its `Border_Closures` schedule on one real locale (as tools/gen_border_golden.py), and
  * two moves appended to that intervention's effect: exposed people leaving the closed locale for the other locales
    (same compartment, other locales: executor.py:185-200; clamped on the first days, when few are exposed) and mild cases
    transferred to hospitals of the other locales (other compartment AND other locales: executor.py:213-233).

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_xmove_golden.py
-> tests/golden/desc_COVID_C_xmove.npz, tests/golden/traj_COVID_C_xmove.npz
"""
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200.export import export_static, export_initial_parameter  # noqa: E402

T_DAYS = 60
EXTRA = ("\n\tsim.move({'compartment':'E', 'locale':locales, 'group':'*'}, {'compartment':'E', 'locale':'~'+locales, 'group':'*'}, cp['degree']*60)"
         "\n\tsim.move({'compartment':'Imild', 'locale':locales, 'group':'*'}, {'compartment':'Hmild', 'locale':'~'+locales, 'group':'*'}, cp['degree']*25)")


def xmove_session():
    s = json.load(open(os.path.join(gg.REF, "jsons", "COVID_C.json")))
    for itv in s["interventions"]:
        if itv["name"] == "Border_Closures":
            itv["effect"] = itv["effect"].rstrip() + EXTRA
    for sch in s["schedules"]:
        if sch["name"] == "Border_Closures":
            sch["detail"][0]["locales"] = "UnitedProvinces.Beaches"
    return s


def main():
    t0 = time.time()
    name = "COVID_C_xmove"
    env = EpiEnv(xmove_session())
    epi = env.epi
    d = export_static(epi)
    d.meta["scenario"] = name
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    out = {f"init_{k[5:]}": v for k, v in export_initial_parameter(epi).items()}
    for k, v in gg.run_schedule_plan(epi, T_DAYS).items():
        out[f"sched_{k}"] = v
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **out)
    print(f"{name}: done in {time.time() - t0:.1f}s  ops={d.NOP} MOVE ops={list(np.asarray(d.op_kind)).count(5)}", flush=True)


if __name__ == "__main__":
    main()
