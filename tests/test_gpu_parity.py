"""Parity of the CUDA path (through the C-ABI) against the reference's own trajectories
(tests/golden) and against the CPU oracle. Tolerances are BASELINE.json's: 1e-9 relative in the
fp64 parity mode, 1e-4 in the fp32 throughput mode (metric: SURVEY.md §8d, absolute floor 1e-6*pop)."""
import numpy as np
import pytest
import torch

from conftest import SCENARIOS
from helpers import cost_err, expand_actions, load, x_err

pytestmark = pytest.mark.gpu

TOL = {"fp64": 1e-9, "fp32": 1e-4}


def _engine(desc, n, prec):
    from epipolicy_rl_b200.engine import BatchedEpidemic
    return BatchedEpidemic(desc, n, device=0, precision=prec)


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_env_plans_per_day_vs_reference(name, prec):
    """rnd and lax weekly plans of the golden file, both stepped day by day in one batch of 2 envs."""
    d, z = load(name)
    eng = _engine(d, 2, prec)
    acts = np.stack([z["rnd_actions"], z["lax_actions"]], 1)  # [53, 2, A]
    X = np.stack([z["rnd_X"], z["lax_X"]], 1)
    C = np.stack([z["rnd_cost"], z["lax_cost"]], 1)
    nf = np.stack([z["rnd_nfev"], z["lax_nfev"]], 1)
    day, worst, worst_c, nfev_mismatch, rew = 0, 0.0, 0.0, 0, np.zeros((len(acts), 2))
    for w, a in enumerate(acts):
        at = torch.from_numpy(a).cuda()
        for _ in range(7):
            if day >= d.horizon:
                break
            _, r, done = eng.step(at, 1)
            rew[w] += r.cpu().numpy()
            worst = max(worst, x_err(eng.get_state("X"), X[day], d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost"), C[day]))
            nfev_mismatch += int(np.any(eng.get_state("trace")[:, 0] != nf[day]))
            day += 1
    assert day == d.horizon and bool(done.all())
    ref_rew = np.stack([z["rnd_rewards"], z["lax_rewards"]], 1)
    rerr = float(np.max(np.abs(rew - ref_rew) / np.abs(ref_rew)))
    print(f"{name} {prec}: X {worst:.2e} cost {worst_c:.2e} reward {rerr:.2e} nfev-mismatch days {nfev_mismatch}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec] and rerr <= TOL[prec]
    assert np.all(np.sign(rew) == np.sign(ref_rew))
    if prec == "fp64":
        assert nfev_mismatch == 0  # identical RK45 accept/reject sequence
    eng.close()


@pytest.mark.parametrize("name", SCENARIOS + ["warmup_scenario"])
def test_schedule_plan_vs_reference(name):
    """BASELINE.json configs[0] (SIR_A fixed plan) and the same for every scenario: the scenario's own
    schedule through daily Epidemic.step."""
    d, z = load(name)
    eng = _engine(d, 1, "fp64")
    X, C = z["sched_X"], z["sched_cost"]
    tot = 0.0
    for day in range(len(X)):
        cpv = torch.from_numpy(z["sched_plan_cpv"][day][None]).cuda()
        act = torch.from_numpy(z["sched_plan_active"][day][None].astype(np.uint8)).cuda()
        _, r, _ = eng.step_plan(cpv, act, True, 1)
        tot += float(r[0])
        assert x_err(eng.get_state("X"), X[day], d) <= 1e-9, day
        assert cost_err(eng.get_state("cost"), C[day]) <= 1e-9, day
        assert int(eng.get_state("trace")[0, 0]) == int(z["sched_nfev"][day])
    assert abs(tot - z["sched_rewards"].sum()) <= 1e-9 * abs(tot)
    if name == "SIR_A":
        assert abs(tot - (-188953003.74567512)) <= 1e-9 * abs(tot)  # SURVEY.md Appendix C anchor
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_border_closure_mobility_multipliers_vs_reference(prec):
    """Mobility multipliers with a non-empty locale selection (core/executor.py:121-134, matrix/coo.py:173-188,
    core_optimize.py:69-92): COVID_C with Border_Closures on one real locale and a change of degree on day 59, a batch of
    3 envs (one warp) following the same schedule, against the unmodified reference (tools/gen_border_golden.py)."""
    d, z = load("COVID_C_border")
    eng = _engine(d, 3, prec)
    assert eng.info.specialized == 1
    X, C = z["sched_X"], z["sched_cost"]
    tol = 1e-9 if prec == "fp64" else 1e-4
    for day in range(len(X)):
        cpv = torch.from_numpy(np.repeat(z["sched_plan_cpv"][day][None], 3, 0)).cuda()
        act = torch.from_numpy(np.repeat(z["sched_plan_active"][day][None].astype(np.uint8), 3, 0)).cuda()
        eng.step_plan(cpv, act, True, 1)
        x = eng.get_state("X")
        for e in range(3):
            assert x_err(x[e], X[day], d) <= tol, (day, e)
            assert cost_err(eng.get_state("cost")[e], C[day]) <= tol, (day, e)
        if prec == "fp64":
            assert (eng.get_state("trace")[:, 0] == int(z["sched_nfev"][day])).all()
    eng.close()


def test_run_batch_of_schedules():
    """BatchedEpidemic.run (Epidemic.run for a batch): env 0 follows SIR_A's own schedule, env 1 the same schedule
    with both degrees raised; env 0 reproduces the golden trajectory, env 1 differs."""
    d, z = load("SIR_A")
    eng = _engine(d, 2, "fp64")
    T = 60
    cpv = torch.from_numpy(z["sched_plan_cpv"][:T]).float()[:, None, :].repeat(1, 2, 1)
    cpv[:, 1, 0], cpv[:, 1, 3] = 0.7, 0.5
    act = torch.from_numpy(z["sched_plan_active"][:T].astype(np.uint8))
    out = eng.run(cpv, act, one_launch=False)  # (the one-launch path: tests/test_zz_gpu_run_plan.py)
    obs = out["obs"].cpu().numpy()
    assert obs.shape == (T, 2, d.C * d.L * d.G) and out["reward"].shape == (T, 2)
    for day in (0, 17, T - 1):
        assert x_err(obs[day, 0], z["sched_X"][day], d) <= 1e-9
    assert np.allclose(out["reward"][:, 0].cpu().numpy(), z["sched_rewards"][:T], rtol=1e-9)
    assert x_err(obs[T - 1, 1], z["sched_X"][T - 1], d) > 1e-6
    eng.close()


def test_known_answer_user_ipynb():
    d, _ = load("warmup_scenario")
    eng = _engine(d, 1, "fp64")
    cpv = d.cp_default.copy()
    cpv[:3] = [0.2, 10, 10]
    _, r, _ = eng.step_plan(torch.from_numpy(cpv[None]).cuda(), None, False, 1)
    assert abs(float(r[0]) - (-4566.504778395923)) <= 1e-9 * 4566.5
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_sirv_b_1024_envs_vs_oracle(prec):
    """BASELINE.json configs[1]: SIRV_B, 1024 envs, env e uses default_rng(1000+e).uniform(0,1,(53,4))."""
    from oracle.oracle import rollout
    d, z = load("SIRV_B")
    n = 1024
    acts = np.stack([np.random.default_rng(1000 + e).uniform(0, 1, (53, d.A)).astype(np.float32) for e in range(n)])
    eng = _engine(d, n, prec)
    rew = np.zeros(n)
    for s in range(53):
        obs, r, done = eng.step(torch.from_numpy(acts[:, s]).cuda(), 7)
        rew += r.cpu().numpy()
    total, fin = rollout(d, acts)
    e = x_err(obs.double().cpu().numpy(), fin, d)
    print(f"SIRV_B x1024 {prec}: final X err {e:.2e}, reward-sum err {abs(rew.sum() - total) / abs(total):.2e}")
    assert e <= TOL[prec]
    assert abs(rew.sum() - total) <= TOL[prec] * abs(total)
    # and the 16 envs the reference itself ran
    assert x_err(obs[:16].double().cpu().numpy(), z["batch_X"][:, -1], d) <= TOL[prec]
    assert np.max(np.abs(rew[:16] - z["batch_rewards"].sum(1)) / np.abs(z["batch_rewards"].sum(1))) <= TOL[prec]
    assert bool(done.all())
    eng.close()


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_covid_c_batch_vs_oracle(prec):
    """BASELINE.json configs[2] spot check: 64 of the 4096 COVID_C envs against the oracle over the FULL horizon
    (53 gym steps, the last one cut at day 365): every gym step's observation and reward of the sampled envs, the
    done flags, and the episode totals."""
    from oracle.oracle import Oracle
    d, _ = load("COVID_C")
    n, steps = 4096, (d.horizon + 6) // 7
    g = torch.Generator().manual_seed(1234)
    acts = torch.rand((steps, n, d.A), generator=g)
    idx = np.arange(0, n, 64)
    oracles = [Oracle(d) for _ in idx]
    eng = _engine(d, n, prec)
    worst = worst_r = 0.0
    for s in range(steps):
        obs, r, done = eng.step(acts[s].cuda(), 7)
        o_gpu, r_gpu, d_gpu = obs[idx].double().cpu().numpy(), r.cpu().numpy()[idx], done.cpu().numpy()[idx]
        for k, (e, o) in enumerate(zip(idx, oracles)):
            o_obs, o_rew, o_done = o.step(acts[s, e].numpy(), 7)
            worst = max(worst, x_err(o_gpu[k], o_obs, d))
            worst_r = max(worst_r, abs(r_gpu[k] - o_rew) / abs(o_rew))
            assert bool(d_gpu[k]) == o_done == (s == steps - 1)
    print(f"COVID_C x4096 {prec}: 64 envs x {steps} gym steps, worst X err {worst:.2e}, worst reward err {worst_r:.2e}")
    assert worst <= TOL[prec] and worst_r <= TOL[prec]
    assert (eng.get_state("day") == d.horizon).all()
    eng.close()


def test_fused_week_equals_seven_days_and_host_path():
    d, _ = load("COVID_B")
    a = torch.rand((16, d.A), generator=torch.Generator().manual_seed(3))
    e1, e7, eh = _engine(d, 16, "fp64"), _engine(d, 16, "fp64"), _engine(d, 16, "fp64")
    r1 = np.zeros(16)
    for _ in range(7):
        _, r, _ = e1.step(a.cuda(), 1)
        r1 += r.cpu().numpy()
    o7, r7, _ = e7.step(a.cuda(), 7)
    oh, rh, dh = eh.step_host(a.numpy(), 7)
    assert np.array_equal(e1.get_state("X"), e7.get_state("X"))  # same arithmetic, bit-identical
    assert np.allclose(r1, r7.cpu().numpy(), rtol=1e-13)
    assert np.array_equal(oh, o7.cpu().numpy()) and np.array_equal(rh, r7.cpu().numpy()) and not dh.any()
    for e in (e1, e7, eh):
        e.close()


@pytest.mark.parametrize("mode", ["0", "1", "3"])
@pytest.mark.parametrize("name,n", [("COVID_C", 100), ("SIR_A", 7), ("syn_s", 5)])
def test_host_path_zero_copy_modes(name, n, mode, monkeypatch):
    """epi_step_host with staged copies (0), results written by the kernel into the mapped pinned buffers (1, the
    default) and the actions read from them as well (3): identical to the device-buffer call, ragged last warp."""
    monkeypatch.setenv("EPI_HOST_ZEROCOPY", mode)
    d, _ = load(name)
    a = torch.rand((n, d.A), generator=torch.Generator().manual_seed(5))
    ed, eh = _engine(d, n, "fp32"), _engine(d, n, "fp32")
    for _ in range(2):
        od, rd, dd = ed.step(a.cuda(), 7)
        oh, rh, dh = eh.step_host(a.numpy(), 7)
        assert np.array_equal(oh, od.cpu().numpy()) and np.array_equal(rh, rd.cpu().numpy())
        assert np.array_equal(dh, dd.cpu().numpy())
    ed.close(); eh.close()


def test_reset_mask_done_and_state_roundtrip():
    d, _ = load("SIRV_A")
    eng = _engine(d, 8, "fp64")
    a = torch.full((8, d.A), 0.3).cuda()
    x0 = eng.get_state("X").copy()
    for _ in range(3):
        eng.step(a, 7)
    assert np.all(eng.get_state("day") == 21)
    mask = torch.tensor([1, 0, 0, 1, 0, 0, 0, 0], dtype=torch.uint8)
    eng.reset(mask)
    day = eng.get_state("day")
    assert list(day) == [0, 21, 21, 0, 21, 21, 21, 21]
    X = eng.get_state("X")
    assert np.array_equal(X[0], x0[0]) and not np.array_equal(X[1], x0[1])
    # a reset env replays exactly the same trajectory
    eng2 = _engine(d, 8, "fp64")
    eng.step(a, 7)
    eng2.step(a, 7)
    assert np.array_equal(eng.get_state("X")[0], eng2.get_state("X")[0])
    # done: horizon reached, stepping continues afterwards like the reference (no auto-reset)
    eng2.set_state("day", np.full(8, d.horizon - 3, np.int32))
    _, _, done = eng2.step(a, 7)
    assert bool(done.all()) and np.all(eng2.get_state("day") == d.horizon)
    _, _, done = eng2.step(a, 7)
    assert bool(done.all()) and np.all(eng2.get_state("day") == d.horizon + 1)
    # flat observable in the reference's layout has the right length and carries X and cost
    flat = eng.get_state("flat")
    assert flat.shape == (8, d.n_flat)
    assert np.array_equal(flat[:, : d.C * d.L * d.G].reshape(X.shape), eng.get_state("X"))
    eng.close()
    eng2.close()


def test_error_paths():
    from epipolicy_rl_b200 import lib
    d, _ = load("SIR_A")
    eng = _engine(d, 4, "fp64")
    with pytest.raises(ValueError):
        eng.step(torch.zeros((3, d.A)).cuda())
    with pytest.raises(lib.EpiError):
        eng.step(torch.zeros((4, d.A)).cuda(), 0)
    with pytest.raises(lib.EpiError):
        eng._ck(eng.L.epi_get_state(eng._h, 0, np.zeros(1).ctypes.data, 8))
    eng.close()


def test_single_env_twin_matches_reference_weekly():
    """EpiEnv twin: float64 numpy obs, python float reward, bool done — checked on the reference's rnd plan."""
    from epipolicy_rl_b200.env import EpiEnv
    d, z = load("COVID_A")
    env = EpiEnv(d)
    obs = env.reset()
    assert obs.dtype == np.float64 and obs.shape == (d.C * d.L * d.G,)
    assert env.action_space.shape == (2,) and env.act_domain.shape == (6, 2)
    for w in range(10):
        obs, r, done, info = env.step(z["rnd_actions"][w])
        assert x_err(obs, z["rnd_obs_weekly"][w], d) <= 1e-9
        assert abs(r - z["rnd_rewards"][w]) <= 1e-9 * abs(r) and done is False and info == {}


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_specialised_kernel_equals_interpreted_kernel(prec, monkeypatch):
    """The scenario-specialised build (spec.py) and the interpreted cell kernel run the same operations in the
    same order: trajectories agree to rounding, RK45 step counts are identical."""
    d, _ = load("COVID_C")
    a = torch.rand((5, 64, d.A), generator=torch.Generator().manual_seed(11))
    es = _engine(d, 64, prec)
    monkeypatch.setenv("EPI_SPEC", "0")
    eg = _engine(d, 64, prec)
    monkeypatch.delenv("EPI_SPEC")
    assert es.info.team_kind == 2 and es.info.specialized == 1 and eg.info.specialized == 0
    for s in range(5):
        o1, r1, _ = es.step(a[s].cuda(), 7)
        o2, r2, _ = eg.step(a[s].cuda(), 7)
        assert np.array_equal(es.get_state("trace"), eg.get_state("trace"))
    tol = 1e-12 if prec == "fp64" else 1e-5
    assert x_err(es.get_state("X"), eg.get_state("X"), d) <= tol
    assert np.allclose(r1.cpu().numpy(), r2.cpu().numpy(), rtol=tol)
    es.close()
    eg.close()


def test_batched_ppo_runs_on_device():
    """The in-repo learner (learner.py) on GPU-resident envs: two PPO updates on COVID_A, everything on the device."""
    from epipolicy_rl_b200.env import BatchedEpiEnv
    from epipolicy_rl_b200.learner import BatchedPPO
    d, _ = load("COVID_A")
    env = BatchedEpiEnv(d, 256, device=0, precision="fp32", auto_reset=False)
    pop = np.broadcast_to(d.pop.reshape(1, d.L, d.G), (d.C, d.L, d.G)).reshape(-1)
    ppo = BatchedPPO(env, env.epi.obs_dim, d.A, obs_scale=pop, n_steps=8, n_epochs=2, device="cuda:0")
    hist = ppo.train(2)
    assert len(hist) == 2 and all(np.isfinite(h["mean_step_reward"]) and h["mean_step_reward"] < 0 for h in hist)
    assert np.isfinite(hist[-1]["v_loss"])
    env.close()


def test_ppo_graph_rollout_equals_eager_rollout_and_gae_on_device():
    """learner numerics on the GPU: (1) the rollout replayed from the captured CUDA graph is the rollout the eager loop
    collects (same seeds, same envs: bit for bit), update after update - which also pins the captured update graph,
    since the rollout of update u is drawn with the weights the updates before it produced; (2) GAE on the device
    against a naive fp64 loop on the host."""
    from epipolicy_rl_b200.env import BatchedEpiEnv
    from epipolicy_rl_b200.learner import BatchedPPO, gae
    d, _ = load("COVID_A")
    pop = np.broadcast_to(d.pop.reshape(1, d.L, d.G), (d.C, d.L, d.G)).reshape(-1)
    runs = []
    for use_graph in (True, False):
        env = BatchedEpiEnv(d, 192, device=0, precision="fp32")
        ppo = BatchedPPO(env, env.epi.obs_dim, d.A, obs_scale=pop, n_steps=6, n_epochs=1, device="cuda:0", seed=3,
                         use_graph=use_graph)
        out = []
        for u in range(3):
            batch = ppo.collect()
            out.append([x.clone() for x in batch])
            ppo.update(batch)
        assert (ppo._graph is not None) == use_graph, "the rollout graph must be in use when asked for"
        assert (ppo._ugraph is not None) == use_graph, "the update graph must be in use when asked for"
        out.append([p.detach().clone() for p in ppo.ac.parameters()])
        runs.append(out)
        env.close()
    for u in range(3):
        for a, b in zip(runs[0][u], runs[1][u]):
            assert torch.equal(a, b), f"update {u}: graph and eager rollouts differ"
    for a, b in zip(runs[0][3], runs[1][3]):  # (3) the weights after three updates (eager, captured, replayed)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), "graph and eager updates differ"
    O, A, LP, V, R, D, last_v = runs[0][2]
    adv, ret = gae(R, V, D, last_v, 0.999, 0.99)
    r64, v64, d64, l64 = (x.double().cpu() for x in (R, V, D, last_v))
    T, N = r64.shape
    ref = torch.zeros((T, N), dtype=torch.float64)
    run, nxt = torch.zeros(N, dtype=torch.float64), l64
    for t in reversed(range(T)):
        nt = 1.0 - d64[t]
        run = (r64[t] + 0.999 * nxt * nt - v64[t]) + 0.999 * 0.99 * nt * run
        ref[t], nxt = run, v64[t]
    assert torch.allclose(adv.double().cpu(), ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_full_size_properties_covid_c_4096(prec):
    """BASELINE.json configs[2] at full size through properties that do not need the oracle:
    (1) people are conserved per (locale, group): sum over compartments == population (deaths are a compartment,
        mobility moves infection pressure, not residents);
    (2) env independence: an env's trajectory does not depend on its position in the batch or on the batch size
        (bit-identical between a 4096-env batch and a 9-env batch fed the same actions);
    (3) permutation: reversing the env order of the actions reverses the results bit for bit."""
    d, _ = load("COVID_C")
    n, steps = 4096, 12
    acts = torch.rand((steps, n, d.A), generator=torch.Generator().manual_seed(1234))
    pick = torch.tensor([0, 1, 2, 3, 1000, 2047, 4093, 4094, 4095])
    big, small, rev = _engine(d, n, prec), _engine(d, len(pick), prec), _engine(d, n, prec)
    for s in range(steps):
        ob, rb, _ = big.step(acts[s].cuda(), 7)
        os_, rs, _ = small.step(acts[s][pick].cuda(), 7)
        orv, rrv, _ = rev.step(acts[s].flip(0).cuda(), 7)
        assert torch.equal(ob[pick.cuda()], os_) and torch.equal(rb[pick.cuda()], rs), s
        assert torch.equal(ob, orv.flip(0)) and torch.equal(rb, rrv.flip(0)), s
    X = big.get_state("X")  # [N, C, L, G] float64
    pop = d.pop.reshape(1, d.L, d.G)
    cons = np.max(np.abs(X.sum(1) - pop) / pop)
    print(f"COVID_C x4096 {prec}: population conservation {cons:.2e}")
    assert cons <= (1e-9 if prec == "fp64" else 1e-5)
    assert np.all(X > -1e-6 * pop[:, None]) and np.all(rb.cpu().numpy() < 0)
    for e in (big, small, rev):
        e.close()
