"""Randomised intervention selectors: every op of a scenario gets random locale / facility / group subsets (so that the
plan compiler produces several locale classes and multiplier classes), facility-interaction multipliers randomly become
facility-timespent multipliers, selects get random masks and value modes, moves random groups and locales - then the
emulated kernels run random weekly actions against the oracle (which is pinned against the reference for every op kind
in test_oracle_golden.py). Deterministic seeds."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from emu_engine import EmuEngine  # noqa: E402
from helpers import cost_err, load, x_err  # noqa: E402

from epipolicy_rl_b200.desc import OP_ADD, OP_FI_MUL, OP_FT_MUL, OP_MOVE, OP_PARAM_MUL  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

TOL = {"fp64": 1e-9, "fp32": 1e-4}


def variant(name, seed):
    rng = np.random.default_rng(seed)
    d, _ = load(name)
    a = d.arrays
    off, data = list(a["list_off"]), list(a["list_data"])

    def newlist(lst):
        data.extend(int(x) for x in lst)
        off.append(len(data))
        return len(off) - 2

    def subset(n):
        return sorted(rng.choice(n, size=rng.integers(1, n + 1), replace=False))

    kind, arg = a["op_kind"].copy(), a["op_arg"].copy().reshape(-1, 8)
    sel = np.asarray(a["sel_arg"]).copy().reshape(-1, 4)
    for o in range(len(kind)):
        if kind[o] == OP_PARAM_MUL:
            if rng.uniform() < 0.7:
                arg[o, 1] = newlist(subset(d.L))
            if rng.uniform() < 0.5:
                arg[o, 2] = newlist(subset(d.F))
            if rng.uniform() < 0.5:
                arg[o, 3] = newlist(subset(d.G))
        elif kind[o] == OP_FI_MUL:
            if rng.uniform() < 0.3:
                kind[o] = OP_FT_MUL
                arg[o, 3] = -1
                if rng.uniform() < 0.5:
                    arg[o, 2] = newlist(subset(d.G))
            else:
                if rng.uniform() < 0.5:
                    arg[o, 2] = newlist(subset(d.G))
                if rng.uniform() < 0.5:
                    arg[o, 3] = newlist(subset(d.G))
            if rng.uniform() < 0.6:
                arg[o, 0] = newlist(subset(d.L))
            if rng.uniform() < 0.4:
                arg[o, 1] = newlist(subset(d.F))
        elif kind[o] == OP_MOVE:
            if rng.uniform() < 0.5:
                g = newlist(subset(d.G))
                arg[o, 2] = arg[o, 5] = g
            if rng.uniform() < 0.4:
                lo = newlist(subset(d.L))
                arg[o, 1] = arg[o, 4] = lo
        elif kind[o] == OP_ADD:
            if rng.uniform() < 0.4:
                arg[o, 1] = newlist(subset(d.L))
    for s in range(len(sel)):
        if rng.uniform() < 0.5:
            sel[s, 1] = newlist(subset(d.L))
        if rng.uniform() < 0.5:
            sel[s, 2] = newlist(subset(d.G))
        # ('change' is left where the scenario has it: over all compartments it is zero up to rounding noise)
        if rng.uniform() < 0.3 and sel[s, 3] == 0:
            sel[s, 3] = 1
    a["list_off"], a["list_data"] = np.array(off, np.int32), np.array(data, np.int32)
    a["op_kind"], a["op_arg"], a["sel_arg"] = kind, arg.reshape(-1), sel.reshape(-1)
    return d.finalize()


CASES = ([("COVID_C", s, "fp64", False, {}) for s in range(6)] + [("COVID_C", s, "fp32", True, {}) for s in range(6)] +
         [("SIRV_B", s, "fp64", True, {}) for s in range(3)] + [("COVID_C", s, "fp64", False, {"EPI_TEAM": "warp"}) for s in range(2)] +
         [("syn_s", s, "fp64", True, {"EPI_GRP": "0"}) for s in range(2)] +
         [("syn_s", s, "fp32", True, {"EPI_GRP": "2", "EPI_GRP_S": "2"}) for s in range(3)])


@pytest.mark.parametrize("name,seed,prec,spec,knobs", CASES,
                         ids=[f"{n}-{s}-{p}-{'spec' if sp else 'interp'}{''.join('-' + v for v in k.values())}" for n, s, p, sp, k in CASES])
def test_random_selectors_vs_oracle(name, seed, prec, spec, knobs, monkeypatch):
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d = variant(name, seed)
    n, weeks = 3, 2
    acts = np.random.default_rng(1000 + seed).uniform(0, 1, (weeks, n, d.A)).astype(np.float32)
    eng = EmuEngine(d, n, prec)
    if spec:
        eng.attach_spec()
    oracles = [Oracle(d) for _ in range(n)]
    worst = worst_c = 0.0
    for w in range(weeks):
        obs, _, _ = eng.step(acts[w], 7)
        for e, o in enumerate(oracles):
            o_obs, _, _ = o.step(acts[w, e], 7)
            worst = max(worst, x_err(obs[e], o_obs, d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost")[e], o.cost))
    print(f"{name} seed {seed} {prec}: kind {eng.info.team_kind} n_lc {eng.info.n_lc} n_cls {eng.info.n_cls} X {worst:.1e} cost {worst_c:.1e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    eng.close()


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("prec,spec", [("fp64", False), ("fp32", True)], ids=["fp64-interp", "fp32-spec"])
def test_random_mobility_multiplier_selections_vs_oracle(seed, prec, spec):
    """the border scenario's mobility multipliers with random modes, groups and from / to locale subsets, stepped through
    the schedule's programs with randomly redrawn control parameters, day by day against the oracle"""
    from epipolicy_rl_b200.desc import OP_MOB_MUL
    rng = np.random.default_rng(500 + seed)
    d, z = load("COVID_C_border")
    a = d.arrays
    off, data = list(a["list_off"]), list(a["list_data"])

    def newlist(n):
        data.extend(int(x) for x in sorted(rng.choice(n, size=rng.integers(1, n + 1), replace=False)))
        off.append(len(data))
        return len(off) - 2

    arg = a["op_arg"].copy().reshape(-1, 8)
    for o in np.where(a["op_kind"] == OP_MOB_MUL)[0]:
        arg[o, :4] = [rng.integers(0, d.M), newlist(d.G), newlist(d.L), newlist(d.L)]
    a["list_off"], a["list_data"], a["op_arg"] = np.array(off, np.int32), np.array(data, np.int32), arg.reshape(-1)
    d.finalize()
    n, T = 3, 12
    cpv = np.repeat(z["sched_plan_cpv"][:T, None, :], n, 1).astype(np.float32)
    act = np.repeat(z["sched_plan_active"][:T, None, :], n, 1).astype(np.uint8)
    free = np.where(d.cp_min != d.cp_max)[0]
    for e in range(n):
        for t in range(T):
            if t == 0 or rng.uniform() < 0.4:
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
    assert worst <= TOL[prec] and worst_c <= TOL[prec], (worst, worst_c)
    eng.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_cross_locale_move_selections_match_or_are_refused(seed):
    """the xmove scenario's two cross-locale moves with random source / target locale subsets and groups: either the
    engine refuses the table (its result would depend on the reference's dictionary order) or the team kernel matches
    the oracle day by day"""
    from epipolicy_rl_b200 import lib
    from epipolicy_rl_b200.desc import OP_MOVE
    rng = np.random.default_rng(700 + seed)
    d, z = load("COVID_C_xmove")
    a = d.arrays
    off, data = list(a["list_off"]), list(a["list_data"])

    def newlist(n, k=None):
        data.extend(int(x) for x in sorted(rng.choice(n, size=k or rng.integers(1, n + 1), replace=False)))
        off.append(len(data))
        return len(off) - 2

    arg = a["op_arg"].copy().reshape(-1, 8)
    for o in np.where(a["op_kind"] == OP_MOVE)[0][-2:]:  # the schedule program's two cross-locale moves
        g = newlist(d.G)
        arg[o, 1], arg[o, 4], arg[o, 2], arg[o, 5] = newlist(d.L), newlist(d.L), g, g
    a["list_off"], a["list_data"], a["op_arg"] = np.array(off, np.int32), np.array(data, np.int32), arg.reshape(-1)
    d.finalize()
    try:
        eng = EmuEngine(d, 2, "fp64")
    except lib.EpiError as e:
        assert e.code == -2
        return
    T = 15
    oracles = [Oracle(d) for _ in range(2)]
    cpv = np.repeat(z["sched_plan_cpv"][:T, None, :], 2, 1).astype(np.float32)
    act = np.repeat(z["sched_plan_active"][:T, None, :], 2, 1).astype(np.uint8)
    free = np.where(d.cp_min != d.cp_max)[0]
    cpv[:, 1, free] = rng.uniform(d.cp_min[free], d.cp_max[free], (T, len(free))).astype(np.float32)
    for t in range(T):
        eng.step_plan(cpv[t], act[t], True, 1)
        for e, o in enumerate(oracles):
            o.day_step(cpv[t, e], act[t, e], init_programs=True)
            assert x_err(eng.get_state("X")[e], o.X, d) <= 1e-9, (t, e)
            assert cost_err(eng.get_state("cost")[e], o.cost) <= 1e-9, (t, e)
    eng.close()
