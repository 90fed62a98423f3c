"""Fresh inputs (not in the golden files): the emulated kernels against the CPU oracle on random weekly actions, every
shipped scenario, both precisions; and the same through the env kernel. The oracle itself is pinned against the
reference in test_oracle_golden.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from conftest import SCENARIOS  # noqa: E402
from emu_engine import EmuEngine  # noqa: E402
from helpers import cost_err, load, x_err  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402

TOL = {"fp64": 1e-9, "fp32": 1e-4}


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_random_weekly_actions_vs_oracle(name, prec):
    d, _ = load(name)
    n, weeks = 4, 6
    rng = np.random.default_rng(sum(name.encode()) + (0 if prec == "fp64" else 7))
    acts = rng.uniform(0, 1, (weeks, n, d.A)).astype(np.float32)
    acts[2, 1] = acts[1, 1]      # a repeated action: the "parameters do not move" path on the first day of a week
    acts[3, 2] = 0.0             # boundary values of the action box
    acts[4, 3] = 1.0
    eng = EmuEngine(d, n, prec)
    oracles = [Oracle(d) for _ in range(n)]
    worst, worst_c, worst_r = 0.0, 0.0, 0.0
    for w in range(weeks):
        obs, rew, done = eng.step(acts[w], 7)
        for e, o in enumerate(oracles):
            o_obs, o_rew, o_done = o.step(acts[w, e], 7)
            worst = max(worst, x_err(obs[e], o_obs, d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost")[e], o.cost))
            worst_r = max(worst_r, abs(rew[e] - o_rew) / abs(o_rew))
            assert bool(done[e]) == bool(o_done)
    print(f"{name} {prec}: X {worst:.2e} cost {worst_c:.2e} reward {worst_r:.2e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec] and worst_r <= TOL[prec]
    eng.close()


@pytest.mark.parametrize("spec", [False, True], ids=["interpreted", "specialised"])
@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_random_border_closures_vs_oracle(prec, spec):
    """mobility multipliers on fresh inputs: three envs of one warp follow different what-if schedules of the border
    scenario (random closure degrees that change on some days and stay on others, so the envs' mobility tables and
    their 'parameters moved today' flags differ inside the warp), day by day against the oracle"""
    d, z = load("COVID_C_border")
    n, T = 3, 24
    rng = np.random.default_rng(17)
    cpv = np.repeat(z["sched_plan_cpv"][:T, None, :], n, 1).astype(np.float32)
    act = np.repeat(z["sched_plan_active"][:T, None, :], n, 1).astype(np.uint8)
    free = np.where(d.cp_min != d.cp_max)[0]
    for e in range(n):
        for t in range(T):
            if t == 0 or rng.uniform() < 0.4:  # redraw every free control parameter (the closure degree among them)
                cpv[t:, e, free] = rng.uniform(d.cp_min[free], d.cp_max[free]).astype(np.float32)
    eng = EmuEngine(d, n, prec)
    if spec:
        eng.attach_spec()
    oracles = [Oracle(d) for _ in range(n)]
    worst = worst_c = 0.0
    for t in range(T):
        eng.step_plan(cpv[t], act[t], True, 1)
        for e, o in enumerate(oracles):
            o.day_step(cpv[t, e], act[t, e], init_programs=True)
            worst = max(worst, x_err(eng.get_state("X")[e], o.X, d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost")[e], o.cost))
    print(f"border random {prec} spec={spec}: X {worst:.2e} cost {worst_c:.2e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_random_cross_locale_moves_vs_oracle(prec):
    """cross-locale moves on fresh inputs: four envs follow different what-if schedules of the xmove scenario (random
    closure degrees, i.e. random amounts asked to leave / to be transferred, redrawn on some days), day by day against
    the oracle; vaccination switched on from day 5 so that same-locale and cross-locale entries are live together"""
    d, z = load("COVID_C_xmove")
    n, T = 4, 20
    rng = np.random.default_rng(23)
    cpv = np.repeat(z["sched_plan_cpv"][:T, None, :], n, 1).astype(np.float32)
    act = np.repeat(z["sched_plan_active"][:T, None, :], n, 1).astype(np.uint8)
    act[5:, :, 0] = 1  # Vaccination has an Act
    free = np.where(d.cp_min != d.cp_max)[0]
    for e in range(n):
        for t in range(T):
            if t == 0 or rng.uniform() < 0.4:
                cpv[t:, e, free] = rng.uniform(d.cp_min[free], d.cp_max[free]).astype(np.float32)
    eng = EmuEngine(d, n, prec)
    oracles = [Oracle(d) for _ in range(n)]
    worst = worst_c = 0.0
    for t in range(T):
        eng.step_plan(cpv[t], act[t], True, 1)
        for e, o in enumerate(oracles):
            o.day_step(cpv[t, e], act[t, e], init_programs=True)
            worst = max(worst, x_err(eng.get_state("X")[e], o.X, d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost")[e], o.cost))
    print(f"xmove random {prec}: X {worst:.2e} cost {worst_c:.2e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_mobility_multipliers_on_a_large_scenario_vs_oracle(prec):
    """L*G > 32 with mobility multipliers: such scenarios are routed to the CTA-team kernel (the env / env-group kernels
    keep static mobility tables). syn_s (20 locales x 4 groups) with one of its parameter multipliers re-targeted at
    the border-mode mobility from locales 0-4 to locales 5-19 (a partial selection, so the row renormalisation does
    not cancel it), random weekly actions against the oracle (which is pinned on mobility multipliers by the border
    golden)."""
    from epipolicy_rl_b200.desc import OP_MOB_MUL, OP_PARAM_MUL
    d, _ = load("syn_s")
    a = d.arrays
    off, data = list(a["list_off"]), list(a["list_data"])
    ids = []
    for lst in (list(range(0, 5)), list(range(5, d.L))):
        data += lst
        off.append(len(data))
        ids.append(len(off) - 2)
    a["list_off"], a["list_data"] = np.array(off, np.int32), np.array(data, np.int32)
    kind, arg = a["op_kind"].copy(), a["op_arg"].copy().reshape(-1, 8)
    o = int(np.where(kind == OP_PARAM_MUL)[0][-1])
    g_list = int(arg[o, 3])
    kind[o] = OP_MOB_MUL
    arg[o] = [1, g_list, ids[0], ids[1], -1, -1, -1, -1]  # mode 1 = border (utility/singleton.py:10)
    a["op_kind"], a["op_arg"] = kind, arg.reshape(-1)
    d.finalize()  # the list pool grew: refresh the descriptor's sizes
    n, weeks = 2, 2
    rng = np.random.default_rng(31)
    acts = rng.uniform(0, 1, (weeks, n, d.A)).astype(np.float32)
    eng = EmuEngine(d, n, prec)
    assert eng.info.team_kind in (0, 1)
    oracles = [Oracle(d) for _ in range(n)]
    plain = Oracle(load("syn_s")[0])
    worst = worst_c = 0.0
    for w in range(weeks):
        obs, rew, _ = eng.step(acts[w], 7)
        p_obs, _, _ = plain.step(acts[w, 0], 7)
        for e, o_ in enumerate(oracles):
            o_obs, o_rew, _ = o_.step(acts[w, e], 7)
            worst = max(worst, x_err(obs[e], o_obs, d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost")[e], o_.cost))
    assert x_err(obs[0], p_obs, d) > 1e-6  # the closure changes the trajectory
    print(f"syn_s + mobility multiplier {prec}: X {worst:.2e} cost {worst_c:.2e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    eng.close()


def _restrict_ops_to_locales(d):
    """interventions restricted to SOME locales (the `locales` regex of a schedule entry, core/executor.py:111-114):
    the first parameter multiplier acts on locale 0 only, the second on locales {1, 2}, the first facility-interaction
    multiplier on {1, 2}, the second on locale 0 - the locales then fall into several LOCALE CLASSES (epi_plan.h), which
    no shipped scenario does"""
    from epipolicy_rl_b200.desc import OP_FI_MUL, OP_PARAM_MUL
    a = d.arrays
    off, data = list(a["list_off"]), list(a["list_data"])
    ids = []
    for lst in ([0], [1, 2]):
        data += lst
        off.append(len(data))
        ids.append(len(off) - 2)
    kind, arg = a["op_kind"], a["op_arg"].copy().reshape(-1, 8)
    pm, fi = np.where(kind == OP_PARAM_MUL)[0], np.where(kind == OP_FI_MUL)[0]
    arg[pm[0], 1] = ids[0]
    if len(pm) > 1:
        arg[pm[1], 1] = ids[1]
    if len(fi) > 0:
        arg[fi[0], 0] = ids[1]
    if len(fi) > 1:
        arg[fi[1], 0] = ids[0]
    a["list_off"], a["list_data"], a["op_arg"] = np.array(off, np.int32), np.array(data, np.int32), arg.reshape(-1)
    return d.finalize()


@pytest.mark.parametrize("name,prec,spec,knobs", [
    ("COVID_C", "fp64", False, {}), ("COVID_C", "fp64", True, {}), ("COVID_C", "fp32", True, {}),
    ("COVID_C", "fp64", False, {"EPI_TEAM": "warp"}), ("syn_s", "fp64", True, {"EPI_GRP": "0"}),
    ("syn_s", "fp32", True, {"EPI_GRP": "2", "EPI_GRP_S": "2"})],
    ids=["cell-interpreted", "cell-specialised", "cell-specialised-fp32", "warp-team", "env-specialised", "env-group-fp32"])
def test_locale_restricted_interventions_vs_oracle(name, prec, spec, knobs, monkeypatch):
    """several locale classes (interventions that act on some locales only) through every kernel family, random weekly
    actions against the oracle"""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d = _restrict_ops_to_locales(load(name)[0])
    n, weeks = 3, 3
    acts = np.random.default_rng(5).uniform(0, 1, (weeks, n, d.A)).astype(np.float32)
    eng = EmuEngine(d, n, prec)
    if spec:
        eng.attach_spec()
    assert eng.info.n_lc > 1
    oracles = [Oracle(d) for _ in range(n)]
    worst = worst_c = 0.0
    for w in range(weeks):
        obs, _, _ = eng.step(acts[w], 7)
        for e, o in enumerate(oracles):
            o_obs, _, _ = o.step(acts[w, e], 7)
            worst = max(worst, x_err(obs[e], o_obs, d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost")[e], o.cost))
    print(f"{name} {prec} n_lc={eng.info.n_lc} kind={eng.info.team_kind}: X {worst:.2e} cost {worst_c:.2e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    eng.close()


@pytest.mark.parametrize("name,prec,spec,knobs", [
    ("COVID_C", "fp64", False, {}), ("COVID_C", "fp64", True, {}), ("COVID_C", "fp32", True, {}),
    ("COVID_C", "fp64", False, {"EPI_TEAM": "warp"}), ("syn_s", "fp32", True, {"EPI_GRP": "2", "EPI_GRP_S": "2"})],
    ids=["cell-interpreted", "cell-specialised", "cell-specialised-fp32", "warp-team", "env-group-fp32"])
def test_move_from_several_compartments_and_some_groups_vs_oracle(name, prec, spec, knobs, monkeypatch):
    """a move whose source selector names several compartments (S and Qs -> V: two move slots fed by one op, the
    departing population summed over both, executor.py:176-183) and only some groups; strong vaccination so that the
    clamp binds when the sources run out"""
    from epipolicy_rl_b200.desc import OP_MOVE
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d = load(name)[0]
    a = d.arrays
    off, data = list(a["list_off"]), list(a["list_data"])
    ids = []
    for lst in ([0, 9], [0, 2]):  # compartments S, Qs; groups 0 and 2
        data += lst
        off.append(len(data))
        ids.append(len(off) - 2)
    arg = a["op_arg"].copy().reshape(-1, 8)
    mv = int(np.where(a["op_kind"] == OP_MOVE)[0][0])
    arg[mv, 0], arg[mv, 2], arg[mv, 5] = ids[0], ids[1], ids[1]
    a["list_off"], a["list_data"], a["op_arg"] = np.array(off, np.int32), np.array(data, np.int32), arg.reshape(-1)
    d.finalize()
    n, weeks = 3, 5
    acts = np.random.default_rng(9).uniform(0, 1, (weeks, n, d.A)).astype(np.float32)
    acts[:, :, 0] = np.maximum(acts[:, :, 0], 0.5)
    eng = EmuEngine(d, n, prec)
    if spec:
        eng.attach_spec()
    assert eng.info.n_slot == 2
    oracles = [Oracle(d) for _ in range(n)]
    worst = worst_c = 0.0
    for w in range(weeks):
        obs, _, _ = eng.step(acts[w], 7)
        for e, o in enumerate(oracles):
            o_obs, _, _ = o.step(acts[w, e], 7)
            worst = max(worst, x_err(obs[e], o_obs, d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost")[e], o.cost))
    print(f"{name} {prec} kind={eng.info.team_kind}: X {worst:.2e} cost {worst_c:.2e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    eng.close()
