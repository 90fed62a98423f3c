"""Golden fixture for the arithmetic of intervention code beyond what the shipped JSONs use (they only multiply and
subtract): addition, division, unary minus, integer and float literals on either side, a select total inside a longer
expression. The reference executes the code as Python over numpy scalars (control parameters are float32, select totals
float64: core/executor.py:272-284,346-359); export.py records it as an expression program with the same promotion and
rounding points, which this fixture pins.

The reference CODE is unmodified; the scenario JSON is COVID_C.json with the `Masks` effect / cost and the `Vaccination`
effect rewritten with such expressions.

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:tools:. python tools/gen_expr_golden.py
-> tests/golden/desc_COVID_C_expr.npz, tests/golden/traj_COVID_C_expr.npz
"""
import json
import os
import time

import numpy as np

import gen_golden as gg  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402
from epipolicy_rl_b200.export import export_static, export_initial_parameter  # noqa: E402

N_WEEKS, T_SCHED = 12, 30


def expr_session():
    s = json.load(open(os.path.join(gg.REF, "jsons", "COVID_C.json")))
    for itv in s["interventions"]:
        if itv["name"] == "Masks":
            itv["effect"] = ("def effect(cp, locales):\n"
                             "\tr = cp['compliance']*cp['max_transmission_reduction']/2 + 0.25*cp['compliance']\n"
                             "\tsim.apply({'parameter':'beta', 'facility':'~Household', 'locale':locales}, 1 - r/(1 + cp['compliance']))\n")
            itv["cost"] = ("def cost(cp, locales):\n"
                           "\tcompliance_count = sim.select({'compartment':'*', 'locale':locales})['Value'].sum() * cp['compliance']\n"
                           "\tsim.add({'locale':locales}, -(-compliance_count)*cp['cost_per_day']/3 + 2 + compliance_count/1000)\n")
        if itv["name"] == "Vaccination":
            itv["effect"] = ("def effect(cp, locales):\n"
                             "\tsim.move({'compartment':'S', 'locale':locales, 'group':'*'}, {'compartment':'V','locale':locales, 'group':'*'}, "
                             "(cp['degree'] + cp['degree']/4)*cp['max_capacity']/1.25)\n")
    return s


def main():
    t0 = time.time()
    name = "COVID_C_expr"
    env = EpiEnv(expr_session())
    epi = env.epi
    d = export_static(epi)
    d.meta["scenario"] = name
    d.save(os.path.join(gg.OUT, f"desc_{name}.npz"))
    out = {f"init_{k[5:]}": v for k, v in export_initial_parameter(epi).items()}
    A = env.action_space.shape[0]
    acts = np.random.default_rng(17).uniform(0, 1, (N_WEEKS, A)).astype(np.float32)
    for k, v in gg.run_env_plan(env, acts).items():
        out[f"rnd_{k}"] = v
    for k, v in gg.run_schedule_plan(epi, T_SCHED).items():
        out[f"sched_{k}"] = v
    np.savez_compressed(os.path.join(gg.OUT, f"traj_{name}.npz"), **out)
    print(f"{name}: done in {time.time() - t0:.1f}s  token ops={sorted(set(int(x) for x in d.arrays['tok_op']))}", flush=True)


if __name__ == "__main__":
    main()
