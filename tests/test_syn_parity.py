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
