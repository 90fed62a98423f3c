"""The env-group kernel (csrc/epi_grp.cuh: one env stepped by S co-resident CTAs, one cell per thread, group barriers)
compiled FOR THE HOST: every CTA of a group runs as lock-stepped host threads with its own shared memory, the S CTAs
run together and meet at the emulated group barrier. Checked against the oracle, the unmodified reference's golden
trajectory and the env kernel (same scenario-specialised program, one CTA per env)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from emu_engine import EmuEngine  # noqa: E402
from helpers import cost_err, load, x_err  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402

TOL = 1e-4  # fp32 throughput mode (the group kernel's only mode)


def _group_engine(d, n, S, monkeypatch):
    monkeypatch.setenv("EPI_GRP", "2")        # also when the env would fit one CTA's shared memory (small test scenario)
    monkeypatch.setenv("EPI_GRP_S", str(S))
    eng = EmuEngine(d, n, "fp32")
    eng.attach_spec()
    assert eng.info.team_kind == 4 and eng.info.reserved == S, (eng.info.team_kind, eng.info.reserved)
    return eng


@pytest.mark.parametrize("S", [1, 2, 3])
def test_group_kernel_vs_oracle_and_env_kernel(S, monkeypatch):
    d, _ = load("syn_s")
    n, weeks = 3, 3
    grp = _group_engine(d, n, S, monkeypatch)
    monkeypatch.setenv("EPI_GRP", "0")
    env = EmuEngine(d, n, "fp32")
    env.attach_spec()
    assert env.info.team_kind == 3
    rng = np.random.default_rng(11 + S)
    acts = rng.uniform(0, 1, (weeks, n, d.A)).astype(np.float32)
    acts[1, 0] = acts[0, 0]  # a repeated action: parameters constant over the week boundary (first-same-as-last path)
    oracles = [Oracle(d) for _ in range(n)]
    worst = worst_c = worst_e = 0.0
    for w in range(weeks):
        obs, rew, done = grp.step(acts[w], 7)
        obs_e, rew_e, _ = env.step(acts[w], 7)
        worst_e = max(worst_e, x_err(obs, obs_e, d))
        assert np.array_equal(grp.get_state("trace")[:, :3], env.get_state("trace")[:, :3])  # same RK45 step sequence
        for e, o in enumerate(oracles):
            o_obs, o_rew, _ = o.step(acts[w, e], 7)
            worst = max(worst, x_err(obs[e], o_obs, d))
            worst_c = max(worst_c, cost_err(grp.get_state("cost")[e], o.cost))
            assert abs(rew[e] - o_rew) <= TOL * abs(o_rew), (w, e, rew[e], o_rew)
        assert not done.any()
    flat_g, flat_e = grp.get_state("flat"), env.get_state("flat")
    acc_err = np.max(np.abs(flat_g - flat_e) / (np.abs(flat_e) + 1e-3))
    print(f"group kernel S={S}: vs oracle X {worst:.2e} cost {worst_c:.2e}; vs env kernel X {worst_e:.2e} flat {acc_err:.2e}")
    assert worst <= TOL and worst_c <= TOL and worst_e <= 1e-5 and acc_err <= 1e-4
    grp.close(); env.close()


def test_group_kernel_vs_reference_per_day(monkeypatch):
    """per-day trajectory of the unmodified reference (tools/gen_syn.py), one launch per simulated day"""
    d, z = load("syn_s")
    grp = _group_engine(d, 1, 2, monkeypatch)
    day, worst, worst_c = 0, 0.0, 0.0
    for a in z["actions"][:2]:
        for _ in range(7):
            grp.step(a[None], 1)
            worst = max(worst, x_err(grp.get_state("X"), z["X"][day], d))
            worst_c = max(worst_c, cost_err(grp.get_state("cost"), z["cost"][day]))
            day += 1
    print(f"group kernel vs reference per day: X {worst:.2e} cost {worst_c:.2e}")
    assert worst <= TOL and worst_c <= TOL
    grp.close()


def test_group_kernel_reset_and_horizon(monkeypatch):
    d, _ = load("syn_s")
    grp = _group_engine(d, 2, 2, monkeypatch)
    a = np.full((2, d.A), 0.3, np.float32)
    o1, r1, _ = (x.copy() for x in grp.step(a, 7))
    grp.reset()
    o2, r2, _ = (x.copy() for x in grp.step(a, 7))
    assert np.array_equal(o1, o2) and np.array_equal(r1, r2)   # deterministic, reset restores everything
    assert np.array_equal(o1[0], o1[1])                         # same action, same env: groups agree bit for bit
    grp.close()


def test_group_kernel_run_plan_one_launch(monkeypatch):
    """schedule mode (epi_run_plan) of the env-group kernel: 6 days of per-day control-parameter rows in ONE launch
    against one epi_step_plan launch per day (fp32: the fused launch reuses a day's last stage as the next day's first
    when the parameters did not move, so agreement is to float rounding, not bit for bit); the per-day reward needs its
    own group sum."""
    d, _ = load("syn_s")
    T, n = 6, 2
    rng = np.random.default_rng(5)
    cpv = np.tile(d.cp_default.astype(np.float32), (T, n, 1))
    free = d.cp_min != d.cp_max
    cpv[:, :, free] = rng.uniform(d.cp_min[free], d.cp_max[free], (T, n, int(free.sum()))).astype(np.float32)
    cpv[3] = cpv[2]  # a day whose parameters do not move
    one, daily = _group_engine(d, n, 2, monkeypatch), _group_engine(d, n, 2, monkeypatch)
    obs_days, rew_days = one.run_plan(cpv, None, False)
    for t in range(T):
        o, r, _ = daily.step_plan(cpv[t], None, False, 1)
        assert x_err(obs_days[t], o, d) <= 1e-5, t
        assert np.allclose(rew_days[t], r, rtol=1e-5), t
        assert x_err(one.x_days[t], daily.get_state("X"), d) <= 1e-5 and np.array_equal(one.mult_days[t], daily.get_state("mult"))
        assert np.allclose(one.acc_days[t], daily.acc(), rtol=1e-4, atol=1e-3), t
    assert x_err(one.get_state("X"), daily.get_state("X"), d) <= 1e-5
    one.close(); daily.close()


@pytest.mark.parametrize("S", [1, 2, 3])
def test_group_kernel_awkward_dimensions_vs_reference(S, monkeypatch):
    """13 locales x 5 groups x 3 facilities (tools/gen_syn_odd.py): a locale's groups straddle warp boundaries, G is neither
    16 nor a multiple of 4 and F is odd, so the mixing pass takes its scalar path and the warps of a CTA have gather rows
    of different lengths - against the unmodified reference's per-day trajectory"""
    d, z = load("syn_odd")
    grp = _group_engine(d, 2, S, monkeypatch)
    day = 0
    for a in z["rnd_actions"]:
        for _ in range(7):
            grp.step(np.repeat(a[None], 2, 0), 1)
            assert x_err(grp.get_state("X")[1], z["rnd_X"][day], d) <= TOL, day
            assert cost_err(grp.get_state("cost")[1], z["rnd_cost"][day]) <= TOL, day
            day += 1
    grp.close()
