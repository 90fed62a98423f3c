"""The scenario-specialised builds (generated per-cell program, prologue tables, constexpr layout) compiled FOR THE
HOST and run by the emulated warps / CTAs: the code the GPU runs in production, checked on the CPU against the oracle
and against the interpreted kernel. (The `-m gpu` tests repeat this on the device.)"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from emu_engine import EmuEngine  # noqa: E402
from helpers import cost_err, load, x_err  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402

TOL = {"fp64": 1e-9, "fp32": 1e-4}


def _run(name, prec, n, weeks, knobs, monkeypatch):
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, _ = load(name)
    rng = np.random.default_rng(sum(name.encode()) + 3)
    acts = rng.uniform(0, 1, (weeks, n, d.A)).astype(np.float32)
    spec, interp = EmuEngine(d, n, prec), EmuEngine(d, n, prec)
    spec.attach_spec()
    assert spec.info.specialized == 1 and interp.info.specialized == 0
    oracles = [Oracle(d) for _ in range(n)]
    worst = worst_c = worst_i = 0.0
    for w in range(weeks):
        obs, rew, _ = spec.step(acts[w], 7)
        obs_i, rew_i, _ = interp.step(acts[w], 7)
        assert np.array_equal(spec.get_state("trace"), interp.get_state("trace"))  # same RK45 step sequence
        worst_i = max(worst_i, x_err(obs, obs_i, d))
        for e, o in enumerate(oracles):
            o_obs, o_rew, _ = o.step(acts[w, e], 7)
            worst = max(worst, x_err(obs[e], o_obs, d))
            worst_c = max(worst_c, cost_err(spec.get_state("cost")[e], o.cost))
            assert abs(rew[e] - o_rew) <= TOL[prec] * abs(o_rew)
    print(f"{name} {prec} {knobs}: vs oracle X {worst:.2e} cost {worst_c:.2e}; vs interpreted {worst_i:.2e}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    assert worst_i <= (1e-12 if prec == "fp64" else 1e-5)
    spec.close(); interp.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("name,n", [("COVID_C", 3), ("SIRV_B", 10), ("SIR_A", 32)])
def test_specialised_cell_kernel_on_the_host(name, n, prec, monkeypatch):
    # n = a full emulated warp of the scenario (32 // (L*G) envs), so the layout is the device's
    _run(name, prec, n, 5, {}, monkeypatch)


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("flowslots", ["0", "1"])
def test_specialised_env_kernel_on_the_host(flowslots, prec, monkeypatch):
    # L*G = 80 > 32: the generated build is the env kernel (one CTA per env), both flow regimes
    _run("syn_s", prec, 2, 2, {"EPI_ENV_FLOWSLOTS": flowslots}, monkeypatch)


def test_forced_kernel_kind_without_a_specialised_build_is_refused(monkeypatch):
    from epipolicy_rl_b200 import lib
    monkeypatch.setenv("EPI_TEAM", "env")
    d, _ = load("COVID_C")
    eng = EmuEngine(d, 2, "fp64")
    with pytest.raises(lib.EpiError) as ei:
        eng.attach_spec()
    assert ei.value.code == -2
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_specialised_cell_kernel_with_mobility_multipliers(prec):
    """the specialised build of a scenario with MOB_MUL ops (per-env mobility tables instead of the dense registers)
    against the unmodified reference's border-closure trajectory"""
    d, z = load("COVID_C_border")
    eng = EmuEngine(d, 1, prec)
    eng.attach_spec()
    assert eng.info.specialized == 1
    X, C = z["sched_X"], z["sched_cost"]
    for day in range(len(X)):
        eng.step_plan(z["sched_plan_cpv"][day][None], z["sched_plan_active"][day][None].astype(np.uint8), True, 1)
        assert x_err(eng.get_state("X"), X[day], d) <= TOL[prec], day
        assert cost_err(eng.get_state("cost"), C[day]) <= TOL[prec], day
        if prec == "fp64":
            assert int(eng.get_state("trace")[0, 0]) == int(z["sched_nfev"][day])
    eng.close()


@pytest.mark.parametrize("prec,spec", [("fp64", False), ("fp64", True), ("fp32", True)], ids=["interpreted", "specialised", "specialised-fp32"])
def test_env_kernel_awkward_dimensions_vs_reference(prec, spec, monkeypatch):
    """13 locales x 5 groups x 3 facilities through the env kernel against the unmodified reference (tools/gen_syn_odd.py)"""
    monkeypatch.setenv("EPI_GRP", "0")
    d, z = load("syn_odd")
    eng = EmuEngine(d, 2, prec)
    if spec:
        eng.attach_spec()
    assert eng.info.team_kind == 3
    day = 0
    for a in z["rnd_actions"]:
        for _ in range(7):
            eng.step(np.repeat(a[None], 2, 0), 1)
            assert x_err(eng.get_state("X")[1], z["rnd_X"][day], d) <= TOL[prec], day
            assert cost_err(eng.get_state("cost")[1], z["rnd_cost"][day]) <= TOL[prec], day
            if prec == "fp64":
                assert int(eng.get_state("trace")[0, 0]) == int(z["rnd_nfev"][day])
            day += 1
    eng.close()


@pytest.mark.parametrize("prec,spec", [("fp64", False), ("fp64", True), ("fp32", True)], ids=["interpreted", "specialised", "specialised-fp32"])
def test_cell_kernel_with_more_than_four_locales_vs_reference(prec, spec):
    """5 locales x 3 groups x 3 facilities (tools/gen_syn_odd.py): the cell kernel's sparse CSC / CSR mobility gathers (its
    dense-register form covers L <= 4, i.e. every shipped scenario) with two envs per warp, four envs = two warps,
    against the unmodified reference"""
    d, z = load("syn_cell")
    eng = EmuEngine(d, 4, prec)
    if spec:
        eng.attach_spec()
    assert eng.info.team_kind == 2 and eng.info.envs_per_cta == 2
    day = 0
    for a in z["rnd_actions"]:
        for _ in range(7):
            eng.step(np.repeat(a[None], 4, 0), 1)
            for e in (0, 3):
                assert x_err(eng.get_state("X")[e], z["rnd_X"][day], d) <= TOL[prec], (day, e)
                assert cost_err(eng.get_state("cost")[e], z["rnd_cost"][day]) <= TOL[prec], (day, e)
            if prec == "fp64":
                assert int(eng.get_state("trace")[3, 0]) == int(z["rnd_nfev"][day])
            day += 1
    eng.close()


@pytest.mark.parametrize("prec,spec", [("fp64", False), ("fp64", True), ("fp32", True)], ids=["interpreted", "specialised", "specialised-fp32"])
def test_cell_kernel_with_a_full_warp_per_env_vs_reference(prec, spec):
    """2 locales x 16 groups x 3 facilities: L*G == 32, the largest scenario of the cell kernel (one env per warp, no idle
    lanes), against the unmodified reference (tools/gen_syn_odd.py)"""
    d, z = load("syn_lg32")
    eng = EmuEngine(d, 2, prec)
    if spec:
        eng.attach_spec()
    assert eng.info.team_kind == 2 and eng.info.envs_per_cta == 1
    day = 0
    for a in z["rnd_actions"]:
        for _ in range(7):
            eng.step(np.repeat(a[None], 2, 0), 1)
            assert x_err(eng.get_state("X")[1], z["rnd_X"][day], d) <= TOL[prec], day
            assert cost_err(eng.get_state("cost")[1], z["rnd_cost"][day]) <= TOL[prec], day
            day += 1
    eng.close()
