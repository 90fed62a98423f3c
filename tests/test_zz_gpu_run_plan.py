"""epi_run_plan on the device: Epidemic.run for a batch of what-if schedules in ONE launch (sorted last: this path was
finished after the round's GPU budget was spent; it has run on the host emulation - tests/test_emu_parity.py::
test_run_plan_one_launch_equals_daily_launches, test_product_host_class_run_on_the_emulation - and these assertions
were dry-run at full length against the emulated library)."""
import os

import numpy as np
import pytest

from helpers import cost_err, load, x_err  # noqa: F401

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]  # a hung kernel must not hang the GPU tier
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(desc, n, prec):
    from epipolicy_rl_b200.engine import BatchedEpidemic
    return BatchedEpidemic(desc, n, device=0, precision=prec)


@pytest.mark.parametrize("name", ["COVID_C", "COVID_C_border"])
def test_run_whole_schedule_in_one_launch_vs_reference(name):
    """epi_run_plan on the device: the scenario's own schedule over its whole golden length for a batch of 48 envs in ONE
    launch - every day's state and reward of every env against the unmodified reference's trajectory (fp64), and the
    same through one launch per day, bit for bit."""
    d, z = load(name)
    T, n = len(z["sched_X"]), 48
    cpv = torch.from_numpy(z["sched_plan_cpv"][:T]).float()
    act = torch.from_numpy(z["sched_plan_active"][:T].astype(np.uint8))
    eng, daily = _engine(d, n, "fp64"), _engine(d, n, "fp64")
    n0 = eng.launches
    out = eng.run(cpv, act)
    assert eng.launches - n0 <= 2
    obs, rew = out["obs"].cpu().numpy(), out["reward"].cpu().numpy()
    for day in range(T):
        assert x_err(obs[day], np.broadcast_to(z["sched_X"][day].reshape(1, -1), (n, d.C * d.L * d.G)), d) <= 1e-9, day
    assert np.allclose(rew, z["sched_rewards"][:T, None], rtol=1e-9)
    daily.reset()
    for day in range(T):
        o, r, _ = daily.step_plan(cpv[day][None].expand(n, -1), act[day][None].expand(n, -1), True, 1)
        assert torch.equal(o, out["obs"][day]) and torch.equal(r, out["reward"][day]), day
    assert bool(out["done"].all()) == (T >= d.horizon)
    eng.close(); daily.close()


@pytest.mark.parametrize("name", ["SIR_A", "COVID_C"])
def test_run_report_in_one_launch_vs_reference_reporter(name):
    """BatchedEpidemic.run(report=True) in one launch: the reporter fed by the per-day emission of epi_run_plan (state in
    double, live cumulative components, class multipliers) against the unmodified reference Reporter, and bit for bit
    against the day-by-day path."""
    d, z = load(name)
    g = np.load(os.path.join(GOLDEN, f"reporter_{name}.npz"))
    T, n = int(g["T"]), 5
    eng = _engine(d, n, "fp64")
    cpv = torch.from_numpy(z["sched_plan_cpv"][:T]).float()[:, None, :].repeat(1, n, 1)
    cpv[:, 1:, :] = cpv[:, 1:, :] * torch.linspace(0.3, 0.9, n - 1)[None, :, None]
    act = torch.from_numpy(z["sched_plan_active"][:T].astype(np.uint8))
    n0 = eng.launches
    one = eng.run(cpv, act, report=True)
    assert eng.launches - n0 <= 2, "reset + ONE step launch"
    daily = eng.run(cpv, act, report=True, one_launch=False)
    rel = lambda a, b, floor: float(np.max(np.abs(a - b) / (np.abs(b) + floor)))  # noqa: E731
    for key, floor in (("average_infectious_duration", 1e-12), ("location_instant_R", 1e-9), ("location_basic_R", 1e-9),
                       ("location_case_R", 1e-9)):
        assert rel(one["report"][key][:, 0].cpu().numpy(), g[key], floor) <= 1e-9, key
        assert torch.equal(one["report"][key], daily["report"][key]), key
    assert torch.equal(one["obs"], daily["obs"]) and torch.equal(one["reward"], daily["reward"])
    eng.close()
