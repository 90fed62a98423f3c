"""Golden fixture for two intervention-code shapes the shipped JSONs never use but the reference
supports: the facility-TIMESPENT multiplier - `sim.apply({'facility', 'locale'}, m)` without group-to, nb_apply's last
branch (core/executor.py:156-162) with the renormalisation over facilities of matrix/dense.py:6-19 - and a select in
value-mode 'default' (core/executor.py:300-303, :346-359).

The reference CODE is unmodified; the scenario JSON is COVID_C.json with the `School closure` effect rewritten in the
timespent form and the `Workplace closure` cost counting the DEFAULT (day-0) adults instead of the current ones.

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_ft_golden.py
-> tests/golden/desc_COVID_C_ft.npz, tests/golden/traj_COVID_C_ft.npz
"""
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200.export import export_static, export_initial_parameter  # noqa: E402

N_WEEKS, T_SCHED = 16, 40


def ft_session():
    s = json.load(open(os.path.join(gg.REF, "jsons", "COVID_C.json")))
    for itv in s["interventions"]:
        if itv["name"] == "School closure":
            itv["effect"] = "def effect(cp, locales):\n    sim.apply({\"facility\": \"School\", \"locale\":locales}, 1-cp['percentage'])\n"
        if itv["name"] == "Workplace closure":
            itv["cost"] = ("def cost(cp, locales):\n    affectedCount = sim.select({'compartment':'*', 'locale':locales, 'group':\"Adults\", "
                           "'value-mode':'default'})['Value'].sum() * cp['percentage']\n"
                           "    sim.add({'locale':locales}, affectedCount*cp['cost_per_day'])\n")
    return s


def main():
    t0 = time.time()
    name = "COVID_C_ft"
    env = EpiEnv(ft_session())
    epi = env.epi
    d = export_static(epi)
    d.meta["scenario"] = name
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    out = {f"init_{k[5:]}": v for k, v in export_initial_parameter(epi).items()}
    A = env.action_space.shape[0]
    acts = np.random.default_rng(11).uniform(0, 1, (N_WEEKS, A)).astype(np.float32)
    for k, v in gg.run_env_plan(env, acts).items():
        out[f"rnd_{k}"] = v
    for k, v in gg.run_schedule_plan(epi, T_SCHED).items():
        out[f"sched_{k}"] = v
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **out)
    kinds = list(np.asarray(d.op_kind))
    sel = np.asarray(d.arrays["sel_arg"]).reshape(-1, 4)
    print(f"{name}: done in {time.time() - t0:.1f}s  ops={d.NOP} FT_MUL ops={kinds.count(3)} select modes={sorted(set(sel[:, 3]))}", flush=True)


if __name__ == "__main__":
    main()
