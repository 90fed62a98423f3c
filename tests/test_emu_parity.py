"""The product's own sources (EpiDesc -> DevPlan compiler, day prologue, RK45 controller, RHS, C-ABI)
compiled for the HOST with a one-thread team (tests/emu) and checked against the reference's golden
trajectories. No GPU needed; the `-m gpu` tests repeat these checks on the real kernels."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from conftest import SCENARIOS  # noqa: E402
from emu_engine import EmuEngine  # noqa: E402
from helpers import cost_err, load, x_err  # noqa: E402

TOL = {"fp64": 1e-9, "fp32": 1e-4}


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_env_plans_per_day_vs_reference(name, prec):
    d, z = load(name)
    eng = EmuEngine(d, 2, prec)
    acts = np.stack([z["rnd_actions"], z["lax_actions"]], 1)
    X = np.stack([z["rnd_X"], z["lax_X"]], 1)
    C = np.stack([z["rnd_cost"], z["lax_cost"]], 1)
    nf = np.stack([z["rnd_nfev"], z["lax_nfev"]], 1)
    day, worst, worst_c, mism, rew = 0, 0.0, 0.0, 0, np.zeros((len(acts), 2))
    for w, a in enumerate(acts):
        for _ in range(7):
            if day >= d.horizon:
                break
            _, r, done = eng.step(a, 1)
            rew[w] += r
            worst = max(worst, x_err(eng.get_state("X"), X[day], d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost"), C[day]))
            mism += int(np.any(eng.get_state("trace")[:, 0] != nf[day]))
            day += 1
    ref_rew = np.stack([z["rnd_rewards"], z["lax_rewards"]], 1)
    rerr = float(np.max(np.abs(rew - ref_rew) / np.abs(ref_rew)))
    print(f"{name} {prec}: X {worst:.2e} cost {worst_c:.2e} reward {rerr:.2e} nfev-mismatch days {mism}")
    assert day == d.horizon and bool(done.all())
    assert worst <= TOL[prec] and worst_c <= TOL[prec] and rerr <= TOL[prec]
    if prec == "fp64":
        assert mism == 0
    eng.close()


@pytest.mark.parametrize("name", SCENARIOS + ["warmup_scenario"])
def test_schedule_plan_vs_reference(name):
    d, z = load(name)
    eng = EmuEngine(d, 1, "fp64")
    X, C = z["sched_X"], z["sched_cost"]
    tot = 0.0
    for day in range(len(X)):
        _, r, _ = eng.step_plan(z["sched_plan_cpv"][day][None], z["sched_plan_active"][day][None].astype(np.uint8), True, 1)
        tot += float(r[0])
        assert x_err(eng.get_state("X"), X[day], d) <= 1e-9, day
        assert cost_err(eng.get_state("cost"), C[day]) <= 1e-9, day
        assert int(eng.get_state("trace")[0, 0]) == int(z["sched_nfev"][day])
    assert abs(tot - z["sched_rewards"].sum()) <= 1e-9 * abs(tot)
    eng.close()


def test_known_answer_and_fused_week():
    d, _ = load("warmup_scenario")
    eng = EmuEngine(d, 1, "fp64")
    cpv = d.cp_default.copy()
    cpv[:3] = [0.2, 10, 10]
    _, r, _ = eng.step_plan(cpv[None], None, False, 1)
    assert abs(float(r[0]) - (-4566.504778395923)) <= 1e-9 * 4566.5
    eng.close()
    d, _ = load("COVID_B")
    a = np.random.default_rng(3).uniform(0, 1, (3, d.A)).astype(np.float32)
    e1, e7 = EmuEngine(d, 3, "fp64"), EmuEngine(d, 3, "fp64")
    r1 = np.zeros(3)
    for _ in range(7):
        r1 += e1.step(a, 1)[1]
    _, r7, _ = e7.step(a, 7)
    assert np.array_equal(e1.get_state("X"), e7.get_state("X")) and np.allclose(r1, r7, rtol=1e-13)


def test_unsupported_programs_are_refused_not_approximated():
    """Moves whose result would depend on the order in which the reference's dictionaries are walked (two moves out of
    one compartment: the clamp of the second sees what the first left, core_optimize.py:196-208) are outside the
    compiled op algebra of the CUDA path: epi_create must fail with EPI_E_UNSUPPORTED rather than run something else."""
    from epipolicy_rl_b200 import lib
    from epipolicy_rl_b200.desc import OP_MOVE
    d, _ = load("COVID_C_xmove")
    kind, arg = d.arrays["op_kind"], d.arrays["op_arg"].copy().reshape(-1, 8)
    mv = np.where(kind == OP_MOVE)[0]
    arg[mv[-1], 0] = arg[mv[-2], 0]  # the hospital transfer now leaves from E as well: (E -> E) and (E -> Hmild)
    d.arrays["op_arg"] = arg.reshape(-1)
    with pytest.raises(lib.EpiError) as ei:
        EmuEngine(d, 1, "fp64")
    assert ei.value.code == -2 and "order-dependent" in str(ei.value)


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("knobs", [{}, {"EPI_TEAM": "cta"}], ids=["warp-team", "cta-team"])
def test_cross_locale_moves_vs_reference(knobs, prec, monkeypatch):
    """Cross-locale moves (nb_move's other-locale branches, core/executor.py:185-200,:213-233): exposed people leaving the
    closed locale for the others (clamped on the first days) and mild cases transferred to the other locales' hospitals,
    against the unmodified reference (tools/gen_xmove_golden.py). Such scenarios run on the team kernels (general move
    entries: the reference's sequential clamp per source cell)."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load("COVID_C_xmove")
    eng = EmuEngine(d, 2, prec)
    assert eng.info.team_kind in (0, 1)
    X, C = z["sched_X"], z["sched_cost"]
    tol = 1e-9 if prec == "fp64" else 1e-4
    for day in range(len(X)):
        cpv = np.repeat(z["sched_plan_cpv"][day][None], 2, 0)
        act = np.repeat(z["sched_plan_active"][day][None].astype(np.uint8), 2, 0)
        eng.step_plan(cpv, act, True, 1)
        for e in range(2):
            assert x_err(eng.get_state("X")[e], X[day], d) <= tol, (day, e)
            assert cost_err(eng.get_state("cost")[e], C[day]) <= tol, (day, e)
        if prec == "fp64":
            assert int(eng.get_state("trace")[0, 0]) == int(z["sched_nfev"][day])
    eng.close()


@pytest.mark.parametrize("knobs", [{}, {"EPI_TEAM": "warp"}, {"EPI_TEAM": "cta"}], ids=["cell", "warp-team", "cta-team"])
def test_border_closure_mobility_multipliers_vs_reference(knobs, monkeypatch):
    """Mobility multipliers with a non-empty locale selection (sim.apply({'locale-from', 'locale-to'}, m),
    core/executor.py:121-134; row renormalisation matrix/coo.py:173-188; per-evaluation combination
    core_optimize.py:69-92): COVID_C with Border_Closures on one real locale and a change of degree on day 59, against
    the unmodified reference's trajectory (tools/gen_border_golden.py), with the identical RK45 step sequence."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load("COVID_C_border")
    eng = EmuEngine(d, 1, "fp64")
    X, C = z["sched_X"], z["sched_cost"]
    for day in range(len(X)):
        eng.step_plan(z["sched_plan_cpv"][day][None], z["sched_plan_active"][day][None].astype(np.uint8), True, 1)
        assert x_err(eng.get_state("X"), X[day], d) <= 1e-9, day
        assert cost_err(eng.get_state("cost"), C[day]) <= 1e-9, day
        assert int(eng.get_state("trace")[0, 0]) == int(z["sched_nfev"][day])
    eng.close()


@pytest.mark.parametrize("knobs", [{"EPI_TEAM": "env", "EPI_ENV_FLOWSLOTS": "0"},
                                   {"EPI_TEAM": "env", "EPI_ENV_FLOWSLOTS": "1"}, {"EPI_TEAM": "warp"}],
                         ids=["env-flowsums", "env-flowslots", "warp-team"])
def test_alternative_kernels_on_covid_c(knobs, monkeypatch):
    """The other kernels of the library on a shipped scenario whose episode includes rejected RK45 attempts:
    the env kernel in both flow regimes, the warp-team kernel."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load("COVID_C")
    eng = EmuEngine(d, 2, "fp64")
    acts = np.stack([z["rnd_actions"], z["lax_actions"]], 1)
    X = np.stack([z["rnd_X"], z["lax_X"]], 1)
    nf = np.stack([z["rnd_nfev"], z["lax_nfev"]], 1)
    day, worst, mism, rej = 0, 0.0, 0, 0
    for a in acts:
        for _ in range(7):
            if day >= d.horizon:
                break
            eng.step(a, 1)
            tr = eng.get_state("trace")
            worst = max(worst, x_err(eng.get_state("X"), X[day], d))
            mism += int(np.any(tr[:, 0] != nf[day]))
            rej += int(tr[:, 2].sum())
            day += 1
    assert worst <= 1e-9 and mism == 0
    assert rej > 0  # the trajectory exercised the reject path
    eng.close()


@pytest.mark.parametrize("name,knobs", [("COVID_C", {}), ("SIRV_B", {}), ("SIR_A", {}),
                                        ("COVID_C", {"EPI_TEAM": "env", "EPI_ENV_FLOWSLOTS": "0"}),
                                        ("COVID_C", {"EPI_TEAM": "env", "EPI_ENV_FLOWSLOTS": "1"})],
                         ids=["COVID_C", "SIRV_B", "SIR_A", "COVID_C-env-flowsums", "COVID_C-env-flowslots"])
def test_fused_week_fp32_reuses_last_stage_across_days(name, knobs, monkeypatch):
    """fp32 throughput mode, 7 fused days: from the second day on f(0, y0) is not evaluated again (the previous
    day's last stage is reused, rk45_day: skip0). Weekly observations stay within the fp32 tolerance of the
    reference, and the RK45 evaluation counts equal those of the day-by-day run that evaluates everything."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load(name)
    fused, daily = EmuEngine(d, 2, "fp32"), EmuEngine(d, 2, "fp32")
    acts = np.stack([z["rnd_actions"], z["lax_actions"]], 1)
    ref = np.stack([z["rnd_obs_weekly"], z["lax_obs_weekly"]], 1)
    worst, weeks = 0.0, min(12, len(ref))
    for w in range(weeks):
        obs, r7, _ = fused.step(acts[w], 7)
        r1 = np.zeros(2)
        for _ in range(7):
            _, r, _ = daily.step(acts[w], 1)
            r1 += r
        worst = max(worst, x_err(obs, ref[w], d))
        assert np.array_equal(fused.get_state("trace")[:, 0], daily.get_state("trace")[:, 0])  # nfev of the last day
        assert np.allclose(r7, r1, rtol=1e-5)
    print(f"{name} fused fp32: X {worst:.2e}")
    assert worst <= TOL["fp32"]
    fused.close(); daily.close()


@pytest.mark.parametrize("knobs", [{}, {"EPI_TEAM": "env"}, {"EPI_TEAM": "warp"}], ids=["cell", "env", "warp-team"])
def test_mirror_outputs_equal_primary_outputs(knobs, monkeypatch):
    """epi_bind_mirror: the second destination (the learner rank's buffer on a GPU box) receives exactly the primary
    results, from the epilogue of the cell / env kernels and from the copy behind the launch for the team kernels;
    ragged last warp (7 envs); unbinding stops the writes."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, _ = load("COVID_C")
    n = 7
    eng = EmuEngine(d, n, "fp32")
    obs2 = np.full((n, d.C * d.L * d.G), -1, np.float32)
    rew2, done2 = np.full(n, -1.0), np.full(n, 9, np.uint8)
    eng._ck(eng.L.epi_bind_mirror(eng._h, obs2.ctypes.data, rew2.ctypes.data, done2.ctypes.data))
    a = np.random.default_rng(4).uniform(0, 1, (n, d.A)).astype(np.float32)
    obs, rew, done = eng.step(a, 7)
    assert np.array_equal(obs2, obs) and np.array_equal(rew2, rew) and np.array_equal(done2, done)
    eng._ck(eng.L.epi_bind_mirror(eng._h, None, None, None))
    keep = obs2.copy()
    eng.step(a, 7)
    assert np.array_equal(obs2, keep) and not np.array_equal(eng.obs, keep)
    eng.close()


def test_create_rejects_bad_arguments():
    """C-ABI error behaviour (integer codes, never throw across the boundary)."""
    import ctypes
    from epipolicy_rl_b200 import lib
    d, _ = load("SIR_A")
    L = EmuEngine(d, 1).L
    c, h = d.as_c(), ctypes.c_void_p()
    assert L.epi_create(None, 1, 0, 0, ctypes.byref(h)) == -1
    assert L.epi_create(ctypes.byref(c), 0, 0, 0, ctypes.byref(h)) == -1   # empty batch
    assert L.epi_create(ctypes.byref(c), -3, 0, 0, ctypes.byref(h)) == -1
    assert L.epi_create(ctypes.byref(c), 1, 0, 7, ctypes.byref(h)) == -1   # unknown precision
    assert L.epi_create(ctypes.byref(c), 1, 0, 0, None) == -1
    assert L.epi_last_error(None)  # a message is available without a handle
    assert L.epi_step(None, None, 7, None, None, None, None) == -1
    assert L.epi_bind_mirror(None, None, None, None) == -1
    with pytest.raises(lib.EpiError):
        e = EmuEngine(d, 2)
        e._ck(L.epi_step(e._h, None, 7, e.obs.ctypes.data, e.reward.ctypes.data, e.done.ctypes.data, None))  # no actions


def test_stepping_past_the_horizon_keeps_simulating():
    """No auto-reset (epidemic.py:116, baselines.py:29-52 steps 100 000 times without reset): done stays set, the day
    counter and the state keep advancing; a fused week stops at the horizon day."""
    d, _ = load("SIR_A")
    eng = EmuEngine(d, 1, "fp64")
    a = np.full((1, d.A), 0.5, np.float32)
    for _ in range(52):
        _, _, done = eng.step(a, 7)
        assert not done[0]
    _, _, done = eng.step(a, 7)  # days 364 (horizon 365): only one of the 7 days runs
    assert done[0] and eng.get_state("day")[0] == 365
    x = eng.get_state("X").copy()
    _, r, done = eng.step(a, 7)
    assert done[0] and eng.get_state("day")[0] > 365 and r[0] < 0
    assert not np.array_equal(eng.get_state("X"), x)
    eng.close()


@pytest.mark.parametrize("name,knobs", [("COVID_C", {}), ("SIR_A", {}), ("COVID_C_border", {}), ("COVID_C", {"EPI_TEAM": "warp"}),
                                        ("COVID_C", {"EPI_TEAM": "env", "EPI_ENV_FLOWSLOTS": "1"})],
                         ids=["COVID_C", "SIR_A", "border", "COVID_C-warp-team", "COVID_C-env"])
def test_run_plan_one_launch_equals_daily_launches(name, knobs, monkeypatch):
    """epi_run_plan (Epidemic.run for a batch, core/epidemic.py:118-134, ONE launch): env 0 follows the scenario's own
    schedule (the reference's trajectory), env 1 the same schedule with its control parameters perturbed from day 20 on;
    per-day states and rewards are bit-identical (fp64) with one epi_step_plan launch per day."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load(name)
    T = 45
    cpv = np.repeat(z["sched_plan_cpv"][:T, None, :], 2, 1).astype(np.float32)
    free = d.cp_min != d.cp_max
    cpv[20:, 1, free] = np.clip(cpv[20:, 1, free] * 1.3 + 0.05, d.cp_min[free], d.cp_max[free])
    act = np.repeat(z["sched_plan_active"][:T, None, :], 2, 1).astype(np.uint8)
    one, daily = EmuEngine(d, 2, "fp64"), EmuEngine(d, 2, "fp64")
    obs_days, rew_days = one.run_plan(cpv, act, True)
    for t in range(T):
        o, r, _ = daily.step_plan(cpv[t], act[t], True, 1)
        assert np.array_equal(obs_days[t], o), t
        assert np.array_equal(rew_days[t], r), t
        # the reporter's inputs, emitted per day: state in double, live cumulative components, class multipliers
        assert np.array_equal(one.x_days[t], daily.get_state("X").reshape(2, -1)), t
        assert np.array_equal(one.acc_days[t], daily.acc()), t
        assert np.array_equal(one.mult_days[t], daily.get_state("mult")), t
    assert x_err(obs_days[:, 0].reshape(T, d.C, d.L, d.G), z["sched_X"][:T], d) <= 1e-9
    assert not np.array_equal(obs_days[-1, 0], obs_days[-1, 1])
    assert np.array_equal(one.get_state("X"), daily.get_state("X")) and np.array_equal(one.get_state("cost"), daily.get_state("cost"))
    one.close(); daily.close()


def test_product_host_class_run_on_the_emulation():
    """engine.BatchedEpidemic.run as shipped (argument handling, EpiRunOut, the reporter fed from the per-day emission),
    executed on the CPU against the emulated library: what-if schedules in one launch; env 0 = the reference's run."""
    import torch
    from host_engine import HostBatchedEpidemic
    d, z = load("COVID_C")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reporter_COVID_C.npz"))
    T, n = int(g["T"]), 3
    eng = HostBatchedEpidemic(d, n, "fp64")
    cpv = torch.from_numpy(z["sched_plan_cpv"][:T]).float()[:, None, :].repeat(1, n, 1)
    cpv[:, 1:, :] = cpv[:, 1:, :] * torch.tensor([0.5, 0.8])[None, :, None]
    act = torch.from_numpy(z["sched_plan_active"][:T].astype(np.uint8))  # [T, I]: broadcast over the envs
    n0 = eng.launches
    out = eng.run(cpv, act, report=True)
    assert eng.launches - n0 <= 2  # reset + ONE step launch
    obs = out["obs"].numpy()
    for day in range(T):
        assert x_err(obs[day, 0], z["sched_X"][day], d) <= 1e-9, day
    assert np.allclose(out["reward"][:, 0].numpy(), z["sched_rewards"][:T], rtol=1e-9)
    assert x_err(obs[T - 1, 1], z["sched_X"][T - 1], d) > 1e-6
    rel = lambda a, b, floor: float(np.max(np.abs(a - b) / (np.abs(b) + floor)))  # noqa: E731
    for key, floor in (("location_instant_R", 1e-9), ("location_basic_R", 1e-9), ("location_case_R", 1e-9)):
        assert rel(out["report"][key][:, 0].numpy(), g[key], floor) <= 1e-9, key
    # the day-by-day path (one epi_step_plan launch per day) gives the same, bit for bit
    daily = eng.run(cpv, act, report=True, one_launch=False)
    assert torch.equal(daily["obs"], out["obs"]) and torch.equal(daily["reward"], out["reward"])
    for key in ("location_instant_R", "location_basic_R", "location_case_R", "incidences"):
        assert torch.equal(daily["report"][key], out["report"][key]), key
    # the [T, NCP] form (one schedule for every env) and no recorded observations
    out2 = eng.run(torch.from_numpy(z["sched_plan_cpv"][:T]), act, record_obs=False)
    assert "obs" not in out2 and torch.equal(out2["reward"][:, 1], out["reward"][:, 0])
    eng.close()


@pytest.mark.parametrize("prec,spec,knobs", [("fp64", False, {}), ("fp64", True, {}), ("fp32", True, {}),
                                             ("fp64", False, {"EPI_TEAM": "warp"}), ("fp64", False, {"EPI_TEAM": "env"})],
                         ids=["cell-interpreted", "cell-specialised", "cell-specialised-fp32", "warp-team", "env"])
def test_facility_timespent_multiplier_vs_reference(prec, spec, knobs, monkeypatch):
    """sim.apply({'facility', 'locale'}, m) (facility-timespent multiplier: the force-of-infection denominator tables
    then depend on the env, `has_ft_ops`) and a value-mode 'default' select, against the unmodified reference
    (tools/gen_ft_golden.py): 16 weekly random actions, day by day"""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load("COVID_C_ft")
    eng = EmuEngine(d, 1, prec)
    if spec:
        eng.attach_spec()
    tol = 1e-9 if prec == "fp64" else 1e-4
    day = 0
    for a in z["rnd_actions"]:
        for _ in range(7):
            eng.step(a[None], 1)
            assert x_err(eng.get_state("X"), z["rnd_X"][day], d) <= tol, day
            assert cost_err(eng.get_state("cost"), z["rnd_cost"][day]) <= tol, day
            if prec == "fp64":
                assert int(eng.get_state("trace")[0, 0]) == int(z["rnd_nfev"][day])
            day += 1
    eng.close()


@pytest.mark.parametrize("prec,spec,knobs", [("fp64", False, {}), ("fp64", True, {}), ("fp32", True, {}),
                                             ("fp64", False, {"EPI_TEAM": "warp"}), ("fp64", False, {"EPI_TEAM": "env"})],
                         ids=["cell-interpreted", "cell-specialised", "cell-specialised-fp32", "warp-team", "env"])
def test_non_conserving_model_vs_reference(prec, spec, knobs, monkeypatch):
    """births and deaths (gain / lost maps, obj/edge.py:76-111; cumulative_gain / cumulative_lost in the flat state and in
    the error norm) against the unmodified reference (tools/gen_vital_golden.py), incl. the flat observable at the end"""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load("SIRV_B_vital")
    eng = EmuEngine(d, 2, prec)
    if spec:
        eng.attach_spec()
    tol = 1e-9 if prec == "fp64" else 1e-4
    day = 0
    for a in z["rnd_actions"]:
        for _ in range(7):
            eng.step(np.repeat(a[None], 2, 0), 1)
            assert x_err(eng.get_state("X")[1], z["rnd_X"][day], d) <= tol, day
            assert cost_err(eng.get_state("cost")[1], z["rnd_cost"][day]) <= tol, day
            if prec == "fp64":
                assert int(eng.get_state("trace")[0, 0]) == int(z["rnd_nfev"][day])
            day += 1
    ref = z["rnd_flat_last"]
    assert np.max(np.abs(eng.get_state("flat")[0] - ref) / (np.abs(ref) + 1e-3)) <= tol
    eng.close()


@pytest.mark.parametrize("prec,spec", [("fp64", False), ("fp64", True), ("fp32", True)],
                         ids=["cell-interpreted", "cell-specialised", "cell-specialised-fp32"])
def test_facility_interaction_multipliers_above_one_vs_reference(prec, spec):
    """facility-interaction rows pushed above a sum of one by their multipliers are rescaled (matrix/dense.py:21-32):
    against the unmodified reference (tools/gen_finorm_golden.py)"""
    d, z = load("COVID_C_finorm")
    eng = EmuEngine(d, 1, prec)
    if spec:
        eng.attach_spec()
    tol = 1e-9 if prec == "fp64" else 1e-4
    for day in range(len(z["sched_X"])):
        eng.step_plan(z["sched_plan_cpv"][day][None], z["sched_plan_active"][day][None].astype(np.uint8), True, 1)
        assert x_err(eng.get_state("X"), z["sched_X"][day], d) <= tol, day
        assert cost_err(eng.get_state("cost"), z["sched_cost"][day]) <= tol, day
        if prec == "fp64":
            assert int(eng.get_state("trace")[0, 0]) == int(z["sched_nfev"][day])
    eng.close()


@pytest.mark.parametrize("prec,spec,knobs", [("fp64", False, {}), ("fp64", True, {}), ("fp32", True, {}), ("fp64", False, {"EPI_TEAM": "warp"})],
                         ids=["cell-interpreted", "cell-specialised", "cell-specialised-fp32", "warp-team"])
def test_intervention_code_arithmetic_vs_reference(prec, spec, knobs, monkeypatch):
    """expression programs with ADD / DIV / NEG tokens and mixed literals (interpreted on the device, generated as
    straight-line code in the specialised build) against the unmodified reference (tools/gen_expr_golden.py)"""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    d, z = load("COVID_C_expr")
    eng = EmuEngine(d, 1, prec)
    if spec:
        eng.attach_spec()
    tol = 1e-9 if prec == "fp64" else 1e-4
    day = 0
    for a in z["rnd_actions"]:
        for _ in range(7):
            eng.step(a[None], 1)
            assert x_err(eng.get_state("X"), z["rnd_X"][day], d) <= tol, day
            assert cost_err(eng.get_state("cost"), z["rnd_cost"][day]) <= tol, day
            if prec == "fp64":
                assert int(eng.get_state("trace")[0, 0]) == int(z["rnd_nfev"][day])
            day += 1
    eng.close()
