"""Cross-locale moves and facility-timespent multipliers on the device (sorted last: this path was finished after the round's GPU budget was spent and has
been run on the host emulation only - tests/test_emu_parity.py::test_cross_locale_moves_vs_reference)."""
import numpy as np
import pytest

from helpers import cost_err, load, x_err

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]  # a hung kernel must not hang the GPU tier


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_cross_locale_moves_vs_reference_on_device(prec):
    """nb_move's other-locale branches (core/executor.py:185-200,:213-233) through the C-ABI on the GPU (team kernel,
    general move entries) against the unmodified reference's trajectory (tools/gen_xmove_golden.py)."""
    from epipolicy_rl_b200.engine import BatchedEpidemic
    d, z = load("COVID_C_xmove")
    n = 5
    eng = BatchedEpidemic(d, n, device=0, precision=prec)
    assert eng.info.team_kind in (0, 1)
    X, C = z["sched_X"], z["sched_cost"]
    tol = 1e-9 if prec == "fp64" else 1e-4
    for day in range(len(X)):
        cpv = torch.from_numpy(np.repeat(z["sched_plan_cpv"][day][None], n, 0)).cuda()
        act = torch.from_numpy(np.repeat(z["sched_plan_active"][day][None].astype(np.uint8), n, 0)).cuda()
        eng.step_plan(cpv, act, True, 1)
        x, c = eng.get_state("X"), eng.get_state("cost")
        for e in range(n):
            assert x_err(x[e], X[day], d) <= tol, (day, e)
            assert cost_err(c[e], C[day]) <= tol, (day, e)
        if prec == "fp64":
            assert (eng.get_state("trace")[:, 0] == int(z["sched_nfev"][day])).all()
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_facility_timespent_multiplier_vs_reference_on_device(prec):
    """sim.apply({'facility', 'locale'}, m) and a value-mode 'default' select (tools/gen_ft_golden.py) on the GPU: the
    fused week of 3 envs of one warp against the unmodified reference's weekly observations and rewards"""
    from epipolicy_rl_b200.engine import BatchedEpidemic
    d, z = load("COVID_C_ft")
    n = 3
    eng = BatchedEpidemic(d, n, device=0, precision=prec)
    tol = 1e-9 if prec == "fp64" else 1e-4
    for w, a in enumerate(z["rnd_actions"]):
        obs, rew, _ = eng.step(torch.from_numpy(np.repeat(a[None], n, 0)).cuda(), 7)
        o = obs.double().cpu().numpy()
        for e in range(n):
            assert x_err(o[e], z["rnd_obs_weekly"][w], d) <= tol, (w, e)
            assert abs(float(rew[e]) - z["rnd_rewards"][w]) <= tol * abs(z["rnd_rewards"][w]), (w, e)
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_non_conserving_model_vs_reference_on_device(prec):
    """births and deaths (gain / lost maps; tools/gen_vital_golden.py) on the GPU: fused weeks of a full warp of envs
    against the unmodified reference's weekly observations, rewards and the flat observable at the end"""
    from epipolicy_rl_b200.engine import BatchedEpidemic
    d, z = load("SIRV_B_vital")
    n = 10
    eng = BatchedEpidemic(d, n, device=0, precision=prec)
    tol = 1e-9 if prec == "fp64" else 1e-4
    for w, a in enumerate(z["rnd_actions"]):
        obs, rew, _ = eng.step(torch.from_numpy(np.repeat(a[None], n, 0)).cuda(), 7)
        o = obs.double().cpu().numpy()
        for e in (0, n - 1):
            assert x_err(o[e], z["rnd_obs_weekly"][w], d) <= tol, (w, e)
            assert abs(float(rew[e]) - z["rnd_rewards"][w]) <= tol * abs(z["rnd_rewards"][w]), (w, e)
    ref = z["rnd_flat_last"]
    assert np.max(np.abs(eng.get_state("flat")[n - 1] - ref) / (np.abs(ref) + 1e-3)) <= tol
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_awkward_dimensions_vs_reference_on_device(prec):
    """13 locales x 5 groups x 3 facilities (tools/gen_syn_odd.py) on the GPU: the env kernel (the scenario fits one CTA's
    shared memory, so it is the default path in both precisions; locales straddle warps, G is not a multiple of 4, F is
    odd) against the unmodified reference's weekly observations and rewards"""
    from epipolicy_rl_b200.engine import BatchedEpidemic
    d, z = load("syn_odd")
    n = 6
    eng = BatchedEpidemic(d, n, device=0, precision=prec)
    assert eng.info.team_kind == 3
    tol = 1e-9 if prec == "fp64" else 1e-4
    for w, a in enumerate(z["rnd_actions"]):
        obs, rew, _ = eng.step(torch.from_numpy(np.repeat(a[None], n, 0)).cuda(), 7)
        o = obs.double().cpu().numpy()
        for e in (0, n - 1):
            assert x_err(o[e], z["rnd_obs_weekly"][w], d) <= tol, (w, e)
            assert abs(float(rew[e]) - z["rnd_rewards"][w]) <= tol * abs(z["rnd_rewards"][w]), (w, e)
    eng.close()


@pytest.mark.parametrize("name", ["syn_cell", "syn_lg32"])
@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_cell_kernel_shape_space_on_device(prec, name):
    """5 x 3 x 3 (sparse mobility gathers, L > 4, two envs per warp) and 2 x 16 x 3 (L*G == 32, a full warp per env)
    (tools/gen_syn_odd.py): the specialised cell kernel on the GPU against the unmodified reference's weekly observations
    and rewards"""
    from epipolicy_rl_b200.engine import BatchedEpidemic
    d, z = load(name)
    n = 7
    eng = BatchedEpidemic(d, n, device=0, precision=prec)
    assert eng.info.team_kind == 2 and eng.info.specialized == 1
    tol = 1e-9 if prec == "fp64" else 1e-4
    for w, a in enumerate(z["rnd_actions"]):
        obs, rew, _ = eng.step(torch.from_numpy(np.repeat(a[None], n, 0)).cuda(), 7)
        o = obs.double().cpu().numpy()
        for e in (0, n - 1):
            assert x_err(o[e], z["rnd_obs_weekly"][w], d) <= tol, (w, e)
            assert abs(float(rew[e]) - z["rnd_rewards"][w]) <= tol * abs(z["rnd_rewards"][w]), (w, e)
    eng.close()
