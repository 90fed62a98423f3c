"""Synthetic COVID-like scenarios (BASELINE.json configs[3] and two smaller sizes; epipolicy_rl_b200/synth.py):
the session is built in the reference's own schema, run through the UNMODIFIED reference (tools/gen_syn.py) and
its per-day trajectory is the golden here. L*G > 32, so these exercise the env kernel (epi_env.cuh: one CTA per env,
threads strided over the cells) and, with EPI_TEAM=cta, the older CTA-team kernel."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from helpers import GOLDEN, cost_err, load, x_err  # noqa: E402

TOL = {"fp64": 1e-9, "fp32": 1e-4}


def _run(eng, d, z, to_dev, n_weeks):
    X, C, nf = z["X"], z["cost"], z["nfev"]
    day, worst, worst_c, mism = 0, 0.0, 0.0, 0
    for a in z["actions"][:n_weeks]:
        for _ in range(7):
            eng.step(to_dev(a[None]), 1)
            worst = max(worst, x_err(eng.get_state("X"), X[day], d))
            worst_c = max(worst_c, cost_err(eng.get_state("cost"), C[day]))
            mism += int(eng.get_state("trace")[0, 0] != nf[day])
            day += 1
    return worst, worst_c, mism


@pytest.mark.parametrize("kernel", ["env", "cta"])
@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("name,weeks", [("syn_s", 2), ("syn_m", 1)])
def test_emulated_kernels_vs_reference(name, weeks, prec, kernel, monkeypatch):
    from emu_engine import EmuEngine
    d, z = load(name)
    if kernel == "cta":
        monkeypatch.setenv("EPI_TEAM", "cta")
    eng = EmuEngine(d, 1, prec)
    assert eng.info.team_kind == (3 if kernel == "env" else 1)
    worst, worst_c, mism = _run(eng, d, z, lambda a: a, weeks)
    print(f"{name} {prec} {kernel} (emulated): X {worst:.2e} cost {worst_c:.2e} nfev mismatches {mism}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    if prec == "fp64":
        assert mism == 0
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("name", ["syn_s", "syn_m", "syn"])
def test_gpu_vs_reference(name, prec):
    import torch
    from epipolicy_rl_b200.engine import BatchedEpidemic
    if not os.path.exists(os.path.join(GOLDEN, f"traj_{name}.npz")):
        pytest.skip("fixture not generated")
    d, z = load(name)
    eng = BatchedEpidemic(d, 1, device=0, precision=prec)
    worst, worst_c, mism = _run(eng, d, z, lambda a: torch.from_numpy(a).cuda(), len(z["actions"]))
    print(f"{name} {prec}: X {worst:.2e} cost {worst_c:.2e} nfev mismatches {mism}")
    assert worst <= TOL[prec] and worst_c <= TOL[prec]
    if prec == "fp64":
        assert mism == 0
    # the weekly reward the reference's EpiEnv.step returned
    eng.reset()
    _, r, _ = eng.step(torch.from_numpy(z["actions"][0][None]).cuda(), 7)
    assert abs(float(r[0]) - z["rewards"][0]) <= TOL[prec] * abs(z["rewards"][0])
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("syn_m", 150), ("syn", 9)])
def test_gpu_group_kernel_vs_env_kernel(name, n, monkeypatch):
    """fp32 throughput mode: the env-group kernel (epi_grp.cuh: S co-resident CTAs per env, cooperative launch,
    working set on chip) against the env kernel (one CTA per env, working set streamed) on the same actions: same
    RK45 step sequences, trajectories within float rounding of each other; more envs than groups on syn (tail)."""
    import torch
    from epipolicy_rl_b200.engine import BatchedEpidemic
    d, _ = load(name)
    grp = BatchedEpidemic(d, n, device=0, precision="fp32")
    assert grp.info.team_kind == 4, "the env-group kernel should be the fp32 path of this scenario"
    monkeypatch.setenv("EPI_GRP", "0")
    env = BatchedEpidemic(d, n, device=0, precision="fp32")
    assert env.info.team_kind == 3
    g = torch.Generator().manual_seed(3)
    worst = 0.0
    for w in range(3):
        a = torch.rand((n, d.A), generator=g).cuda()
        if w == 1:
            a[0] = a_prev[0]  # repeated action: first-same-as-last across the week boundary
        a_prev = a
        og, rg, dg = grp.step(a, 7)
        oe, re_, de = env.step(a, 7)
        torch.cuda.synchronize()
        assert np.array_equal(grp.get_state("trace")[:, :3], env.get_state("trace")[:, :3])
        worst = max(worst, x_err(og.double().cpu().numpy(), oe.double().cpu().numpy(), d))
        assert torch.allclose(rg, re_, rtol=1e-5)
        assert torch.equal(dg, de)
    flat_g, flat_e = grp.get_state("flat"), env.get_state("flat")
    acc_err = float(np.max(np.abs(flat_g - flat_e) / (np.abs(flat_e) + 1e-3)))
    print(f"{name}: group vs env kernel X {worst:.2e} flat {acc_err:.2e}; grid {grp.info.grid} x {grp.info.threads}, "
          f"S {grp.info.reserved}, smem {grp.info.smem_bytes}")
    assert worst <= 1e-5 and acc_err <= 1e-4
    # a masked reset and a second episode start reproduce the first step bit for bit
    grp.reset()
    a = torch.rand((n, d.A), generator=torch.Generator().manual_seed(3)).cuda()
    o2, _, _ = grp.step(a, 7)
    env.reset()
    o3, _, _ = env.step(a, 7)
    assert x_err(o2.double().cpu().numpy(), o3.double().cpu().numpy(), d) <= 1e-5
    grp.close(); env.close()
