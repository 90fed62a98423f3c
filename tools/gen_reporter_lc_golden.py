"""Reporter golden for a scenario with SEVERAL LOCALE CLASSES: COVID_C with `Masks` scheduled on one locale and
`Mass_Screening_and_Contact_Tracing` on another (the shipped schedules act on all locales, so every shipped scenario
collapses to one class). The reference CODE is unmodified. The reference Reporter's R-number chain
(core/reporter.py:54-204) averages the model parameters of each reporting location's locales - the device reporter has to
expand its per-class tables by locale to reproduce that.

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_reporter_lc_golden.py
-> tests/golden/desc_COVID_C_lc.npz, traj_COVID_C_lc.npz, reporter_COVID_C_lc.npz
"""
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200.export import export_static, reporter_meta  # noqa: E402

T = 28


def lc_session():
    s = json.load(open(os.path.join(gg.REF, "jsons", "COVID_C.json")))
    for sch in s["schedules"]:
        d0 = sch["detail"][0]
        if sch["name"] == "Masks":
            d0["locales"] = "UnitedProvinces.Beaches"
            d0["start_date"] = "2021-01-01"
            d0["control_params"][0]["value"] = "0.7"
        if sch["name"] == "Mass_Screening_and_Contact_Tracing":
            d0["locales"] = "UnitedProvinces.Hills"
    return s


def main():
    t0 = time.time()
    name = "COVID_C_lc"
    session = lc_session()
    env = EpiEnv(session)
    epi = env.epi
    d = export_static(epi)
    d.meta["scenario"] = name
    d.meta["reporter"] = reporter_meta(epi.static, session)
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    out = {f"sched_{k}": v for k, v in gg.run_schedule_plan(epi, T).items()}
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **out)
    rep = epi.reporter
    rep.compute_history(); rep.compute_incidences(); rep.compute_comp_news()
    rep.compute_average_infectious_duration(); rep.compute_location_statistics(); rep.compute_window_statistics()
    rep.compute_instantaneous_reproductive_number(); rep.compute_basic_reproductive_number(); rep.compute_case_reproductive_number()
    np.savez_compressed(os.path.join(gg.OUT, f"reporter_{name}.npz"), incidences=np.asarray(rep.incidences), T=T,
                        average_infectious_duration=np.asarray(rep.average_infectious_duration),
                        window_infectious_duration=np.asarray(rep.window_infectious_duration),
                        location_window_incidence=np.asarray(rep.location_window_incidence),
                        location_window_infectious_count=np.asarray(rep.location_window_infectious_count),
                        location_instant_R=np.asarray(rep.location_instant_R),
                        location_basic_R=np.asarray(rep.location_basic_R),
                        location_case_R=np.asarray(rep.location_case_R))
    print(f"{name}: done in {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
