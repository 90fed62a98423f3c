"""The CPU oracle against outputs of the unmodified reference (tests/golden, made by tools/gen_golden.py).
This is what pins the oracle (the reference ships no tests of its own)."""
import numpy as np
import pytest

from conftest import SCENARIOS
from helpers import cost_err, load, x_err
from oracle.oracle import Oracle, rollout


@pytest.mark.parametrize("name", SCENARIOS + ["warmup_scenario"])
def test_initial_parameter_literal(name):
    d, z = load(name)
    p = Oracle(d).params(1)
    for k in ("mp", "ft", "fi", "mob"):
        assert np.array_equal(p[k], z[f"init_{k}"].reshape(p[k].shape)), k  # float32 tables: bit-exact
    assert np.allclose(p["cost"], z["init_cost"], rtol=1e-12, atol=0)
    assert np.array_equal(p["moves"], z["init_moves"].reshape(p["moves"].shape))


@pytest.mark.parametrize("name", SCENARIOS)
@pytest.mark.parametrize("plan", ["rnd", "lax"])
def test_env_plans_per_day(name, plan):
    d, z = load(name)
    o = Oracle(d)
    X, C, acts = z[f"{plan}_X"], z[f"{plan}_cost"], z[f"{plan}_actions"]
    day, nfev, rew = 0, [], []
    for a in acts:
        tot = 0.0
        for _ in range(7):
            if day >= d.horizon:
                break
            _, r, _ = o.step(a, 1)
            tot += r
            assert x_err(o.X, X[day], d) <= 1e-11, (day,)
            assert cost_err(o.cost, C[day]) <= 1e-11, (day,)
            if day < 10:  # the day's dest parameters (float32 tables bit-exact)
                p = o.params(0)
                for k in ("mp", "ft", "fi", "mob"):
                    assert np.array_equal(p[k], z[f"{plan}_p_{k}"][day].reshape(p[k].shape)), (day, k)
                assert np.array_equal(p["moves"], z[f"{plan}_p_moves{day}"].reshape(p["moves"].shape)), day
                assert np.allclose(p["cost"], z[f"{plan}_p_cost"][day], rtol=1e-11, atol=1e-6)
            nfev.append(o.trace["nfev"])
            day += 1
        rew.append(tot)
    assert np.array_equal(np.array(nfev), z[f"{plan}_nfev"])  # identical RK45 step sequence
    assert np.allclose(rew, z[f"{plan}_rewards"], rtol=1e-11)


@pytest.mark.parametrize("name", SCENARIOS + ["warmup_scenario"])
def test_schedule_plan(name):
    d, z = load(name)
    o = Oracle(d)
    X, C = z["sched_X"], z["sched_cost"]
    rew, nfev = [], []
    for day in range(len(X)):
        rew.append(o.day_step(z["sched_plan_cpv"][day], z["sched_plan_active"][day], init_programs=True))
        assert x_err(o.X, X[day], d) <= 1e-11
        assert cost_err(o.cost, C[day]) <= 1e-11
        nfev.append(o.trace["nfev"])
    assert np.array_equal(np.array(nfev), z["sched_nfev"])
    assert np.allclose(rew, z["sched_rewards"], rtol=1e-10, atol=1e-6)


def test_known_answer_user_ipynb():
    """user.ipynb:82 — epi.step([construct_act(0,[0.2,10,10])]) on the warm-up scenario."""
    d, z = load("warmup_scenario")
    o = Oracle(d)
    cpv = d.cp_default.copy()
    cpv[:3] = [0.2, 10, 10]
    r = o.day_step(cpv, np.ones(d.I, np.int32))
    assert abs(r - (-4566.504778395923)) <= 1e-9
    assert abs(float(z["known_answer_reward"][0]) - (-4566.504778395923)) == 0.0


def test_sirv_b_batch_subset():
    """BASELINE.json configs[1]: env e uses default_rng(1000+e).uniform(0,1,(53,4))."""
    d, z = load("SIRV_B")
    n = z["batch_X"].shape[0]
    acts = np.stack([np.random.default_rng(1000 + e).uniform(0, 1, (53, d.A)).astype(np.float32) for e in range(n)])
    total, fin = rollout(d, acts)
    assert x_err(fin, z["batch_X"][:, -1], d) <= 1e-11
    assert abs(total - z["batch_rewards"].sum()) <= 1e-11 * abs(total)


def test_schedule_plan_border_closure_mobility_multipliers():
    """Mobility multipliers with a NON-EMPTY locale selection (core/executor.py:121-134, matrix/coo.py:112-117,173-188):
    COVID_C with its Border_Closures schedule pointed at one real locale and a change of degree on day 59, run through
    the unmodified reference (tools/gen_border_golden.py). The mode matrices of the first ten days bit-exact, the
    90-day trajectory and costs, the RK45 step sequence."""
    d, z = load("COVID_C_border")
    assert list(d.op_kind).count(4) == 2  # the two sim.apply({'locale-from', 'locale-to'}) of the effect code
    o = Oracle(d)
    p = o.params(1)
    for k in ("mp", "ft", "fi", "mob"):
        assert np.array_equal(p[k], z[f"init_{k}"].reshape(p[k].shape)), k
    X, C = z["sched_X"], z["sched_cost"]
    nfev = []
    for day in range(len(X)):
        o.day_step(z["sched_plan_cpv"][day], z["sched_plan_active"][day], init_programs=True)
        assert x_err(o.X, X[day], d) <= 1e-11
        assert cost_err(o.cost, C[day]) <= 1e-11
        if day < 10:
            p = o.params(0)
            assert np.array_equal(p["mob"], z["sched_p_mob"][day].reshape(p["mob"].shape)), day
        nfev.append(o.trace["nfev"])
    assert np.array_equal(np.array(nfev), z["sched_nfev"])
    # the closure changes the dynamics: the same schedule without it (all other rows equal) ends somewhere else
    zc = load("COVID_C")[1]
    assert x_err(X[59], zc["sched_X"][59], d) > 1e-4


def test_schedule_plan_cross_locale_moves():
    """Cross-locale moves (nb_move's other-locale branches, core/executor.py:185-200,:213-233; sequential clamp per source
    cell, core_optimize.py:196-208), which no shipped intervention uses: COVID_C with two such moves appended to the
    Border_Closures effect and the schedule on one locale, through the unmodified reference (tools/gen_xmove_golden.py).
    The move tables of the first ten days (same entries and float32 amounts; the reference lists them group by group,
    the oracle in insertion order), the 60-day trajectory and costs, the RK45 step sequence."""
    d, z = load("COVID_C_xmove")
    o = Oracle(d)
    X, C = z["sched_X"], z["sched_cost"]
    nfev, clamped = [], 0
    for day in range(len(X)):
        o.day_step(z["sched_plan_cpv"][day], z["sched_plan_active"][day], init_programs=True)
        assert x_err(o.X, X[day], d) <= 1e-11
        assert cost_err(o.cost, C[day]) <= 1e-11
        if day < 10:
            mine, ref = o.params(0)["moves"], z[f"sched_p_moves{day}"].reshape(-1, 6)
            key = lambda m: m[np.lexsort(m[:, :5].T[::-1])]  # noqa: E731
            assert np.array_equal(key(mine), key(ref)), day
            if len(ref):
                cross = ref[ref[:, 3] != ref[:, 4]]
                assert len(cross) > 0
                E0 = X[day - 1].reshape(d.C, d.L, d.G)[1, 0] if day else None  # exposed in the closed locale, yesterday
                if E0 is not None:
                    clamped += int(np.any(cross[cross[:, 1] == 1][:, 5].reshape(d.G, -1).sum(1) > E0))
        nfev.append(o.trace["nfev"])
    assert np.array_equal(np.array(nfev), z["sched_nfev"])
    assert clamped > 0  # on the first days more exposed people are asked to leave than there are: the clamp binds


def test_facility_timespent_multiplier_and_default_mode_select():
    """Two intervention-code shapes no shipped JSON uses: sim.apply({'facility', 'locale'}, m) - the facility-timespent
    multiplier of nb_apply's last branch (core/executor.py:156-162), renormalised over the facilities
    (matrix/dense.py:6-19) - and a select with value-mode 'default' (core/executor.py:300-303). COVID_C with School
    closure rewritten in that form, through the unmodified reference (tools/gen_ft_golden.py): weekly random actions
    through EpiEnv.step and the schedule through Epidemic.step; float32 tables of the first ten days bit-exact."""
    d, z = load("COVID_C_ft")
    assert list(d.op_kind).count(3) == 1 and 1 in set(np.asarray(d.arrays["sel_arg"]).reshape(-1, 4)[:, 3])
    o = Oracle(d)
    p = o.params(1)
    for k in ("mp", "ft", "fi", "mob"):
        assert np.array_equal(p[k], z[f"init_{k}"].reshape(p[k].shape)), k
    X, C = z["rnd_X"], z["rnd_cost"]
    day, nfev = 0, []
    for a in z["rnd_actions"]:
        for _ in range(7):
            o.step(a, 1)
            assert x_err(o.X, X[day], d) <= 1e-11 and cost_err(o.cost, C[day]) <= 1e-11, day
            if day < 10:
                p = o.params(0)
                for k in ("mp", "ft", "fi"):
                    assert np.array_equal(p[k], z[f"rnd_p_{k}"][day].reshape(p[k].shape)), (day, k)
            nfev.append(o.trace["nfev"])
            day += 1
    assert np.array_equal(np.array(nfev), z["rnd_nfev"])
    o = Oracle(d)
    for day in range(len(z["sched_X"])):
        o.day_step(z["sched_plan_cpv"][day], z["sched_plan_active"][day], init_programs=True)
        assert x_err(o.X, z["sched_X"][day], d) <= 1e-11 and cost_err(o.cost, z["sched_cost"][day]) <= 1e-11, day


def test_non_conserving_model_gain_and_lost_maps():
    """Equation terms without a counterpart (births into S, deaths out of R) become gain / lost entries of their edges
    (obj/edge.py:76-111) and feed cumulative_gain / cumulative_lost (core_optimize.py:171-195), which sit in the flat
    ODE state and in the RK45 error norm; every shipped model conserves mass. SIRV_B with vital dynamics through the
    unmodified reference (tools/gen_vital_golden.py): 140 days of weekly random actions, the flat observable at the end."""
    d, z = load("SIRV_B_vital")
    assert {1, 2} <= set(int(k) for k in d.arrays["sc_kind"])  # EPI_SC_GAIN, EPI_SC_LOST
    o = Oracle(d)
    day, nfev = 0, []
    for a in z["rnd_actions"]:
        for _ in range(7):
            o.step(a, 1)
            assert x_err(o.X, z["rnd_X"][day], d) <= 1e-11 and cost_err(o.cost, z["rnd_cost"][day]) <= 1e-11, day
            nfev.append(o.trace["nfev"])
            day += 1
    assert np.array_equal(np.array(nfev), z["rnd_nfev"])
    ref = z["rnd_flat_last"]
    assert np.max(np.abs(o.flat - ref) / (np.abs(ref) + 1e-3)) <= 1e-11
    assert z["rnd_X"][-1].sum() > 1.5 * z["rnd_X"][0].sum()  # the population grows: mass is not conserved


def test_facility_interaction_rows_renormalised_under_multipliers_above_one():
    """get_normalized_facility_interaction (matrix/dense.py:21-32) rescales a row only when its sum is > 1 or < 0; with the
    shipped schedules (multipliers <= 1) only the default rows reach that branch. COVID_C with negative closure
    percentages (multipliers 1.6 / 1.3, other values from day 20) through the unmodified reference
    (tools/gen_finorm_golden.py): the float32 tables of the first ten days bit-exact, 45 days of trajectory and costs."""
    d, z = load("COVID_C_finorm")
    o = Oracle(d)
    nfev = []
    for day in range(len(z["sched_X"])):
        o.day_step(z["sched_plan_cpv"][day], z["sched_plan_active"][day], init_programs=True)
        assert x_err(o.X, z["sched_X"][day], d) <= 1e-11 and cost_err(o.cost, z["sched_cost"][day]) <= 1e-11, day
        if day < 10:
            p = o.params(0)
            for k in ("mp", "ft", "fi", "mob"):
                assert np.array_equal(p[k], z[f"sched_p_{k}"][day].reshape(p[k].shape)), (day, k)
        nfev.append(o.trace["nfev"])
    assert np.array_equal(np.array(nfev), z["sched_nfev"])
    assert float(z["sched_plan_cpv"][0].min()) < 0  # the multipliers 1 - percentage are above one


def test_intervention_code_arithmetic_beyond_the_shipped_shapes():
    """Addition, division, unary minus, integer / float literals and a select total inside longer expressions (the shipped
    JSONs only multiply and subtract): the recorder's numpy-scalar promotion and rounding points (export.py) and the
    expression VM against the unmodified reference executing the Python code (tools/gen_expr_golden.py): float32
    parameter tables and move tables of the first ten days bit-exact, cost rates to 1e-12, 84 days of trajectory."""
    d, z = load("COVID_C_expr")
    assert {3, 6, 7} <= set(int(x) for x in d.arrays["tok_op"])  # ADD, DIV, NEG
    o = Oracle(d)
    day, nfev = 0, []
    for a in z["rnd_actions"]:
        for _ in range(7):
            o.step(a, 1)
            assert x_err(o.X, z["rnd_X"][day], d) <= 1e-11 and cost_err(o.cost, z["rnd_cost"][day]) <= 1e-11, day
            if day < 10:
                p = o.params(0)
                for k in ("mp", "ft", "fi", "mob"):
                    assert np.array_equal(p[k], z[f"rnd_p_{k}"][day].reshape(p[k].shape)), (day, k)
                assert np.array_equal(p["moves"], z[f"rnd_p_moves{day}"].reshape(p["moves"].shape)), day
                assert np.allclose(p["cost"], z["rnd_p_cost"][day], rtol=1e-12, atol=0), day
            nfev.append(o.trace["nfev"])
            day += 1
    assert np.array_equal(np.array(nfev), z["rnd_nfev"])
    o = Oracle(d)
    for day in range(len(z["sched_X"])):
        o.day_step(z["sched_plan_cpv"][day], z["sched_plan_active"][day], init_programs=True)
        assert x_err(o.X, z["sched_X"][day], d) <= 1e-11 and cost_err(o.cost, z["sched_cost"][day]) <= 1e-11, day


def test_schedule_on_single_locales_several_locale_classes():
    """interventions scheduled on single locales (tools/gen_reporter_lc_golden.py): the oracle against the unmodified
    reference, float32 tables of the first ten days bit-exact"""
    d, z = load("COVID_C_lc")
    o = Oracle(d)
    nfev = []
    for day in range(len(z["sched_X"])):
        o.day_step(z["sched_plan_cpv"][day], z["sched_plan_active"][day], init_programs=True)
        assert x_err(o.X, z["sched_X"][day], d) <= 1e-11 and cost_err(o.cost, z["sched_cost"][day]) <= 1e-11, day
        if day < 10:
            p = o.params(0)
            for k in ("mp", "ft", "fi", "mob"):
                assert np.array_equal(p[k], z[f"sched_p_{k}"][day].reshape(p[k].shape)), (day, k)
        nfev.append(o.trace["nfev"])
    assert np.array_equal(np.array(nfev), z["sched_nfev"])
    mp = z["sched_p_mp"][0]
    assert not np.array_equal(mp[:, 0], mp[:, 1])  # the locales' parameters differ: masks in one, screening in another


def test_synthetic_scenario_with_awkward_dimensions():
    """13 locales x 5 groups x 3 facilities (tools/gen_syn_odd.py): the oracle against the unmodified reference"""
    d, z = load("syn_odd")
    assert (d.L, d.G, d.F) == (13, 5, 3)
    o = Oracle(d)
    day, nfev = 0, []
    for a in z["rnd_actions"]:
        for _ in range(7):
            o.step(a, 1)
            assert x_err(o.X, z["rnd_X"][day], d) <= 1e-11 and cost_err(o.cost, z["rnd_cost"][day]) <= 1e-11, day
            nfev.append(o.trace["nfev"])
            day += 1
    assert np.array_equal(np.array(nfev), z["rnd_nfev"])


@pytest.mark.parametrize("name,dims", [("syn_cell", (5, 3, 3)), ("syn_lg32", (2, 16, 3))])
def test_synthetic_scenario_in_the_cell_kernels_shape_space(name, dims):
    """5 x 3 x 3 (L*G <= 32 with more than four locales) and 2 x 16 x 3 (L*G == 32: a full warp per env, the boundary of
    the cell kernel) through the unmodified reference (tools/gen_syn_odd.py)"""
    d, z = load(name)
    assert (d.L, d.G, d.F) == dims
    o = Oracle(d)
    day, nfev = 0, []
    for a in z["rnd_actions"]:
        for _ in range(7):
            o.step(a, 1)
            assert x_err(o.X, z["rnd_X"][day], d) <= 1e-11 and cost_err(o.cost, z["rnd_cost"][day]) <= 1e-11, day
            nfev.append(o.trace["nfev"])
            day += 1
    assert np.array_equal(np.array(nfev), z["rnd_nfev"])
