"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported via PYTHONPATH) in this container. The GPU box has no reference,
so these committed fixtures are what pins the oracle and the CUDA path.

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:. python tools/gen_golden.py [names...]

For each scenario: desc_<name>.npz (EpiDesc), traj_<name>.npz (per-day trajectories of three
plans + the detached initial Parameter literal + a few days of dest parameters + RK45 nfev).
"""
import json
import os
import sys
import time

import numpy as np

REF = os.environ.get("EPI_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
SCN = os.path.join(os.path.dirname(__file__), "..", "epipolicy_rl_b200", "scenarios")  # product data: the descriptors

import epipolicy.core.epidemic as ref_epidemic  # noqa: E402
from epipolicy.obj.act import construct_act  # noqa: E402
from epipolicy_environment import EpiEnv  # noqa: E402

from epipolicy_rl_b200.export import export_static, export_initial_parameter  # noqa: E402

_trace = []
_orig_solve_ivp = ref_epidemic.solve_ivp


def _traced(*a, **k):
    res = _orig_solve_ivp(*a, **k)
    _trace.append((res.nfev, len(res.t) - 1, res.status))
    return res


ref_epidemic.solve_ivp = _traced


def moves_of(param):
    out = []
    d3 = param.delta_comp_to_comp
    for g in d3:
        for c1 in d3[g]:
            for c2 in d3[g][c1]:
                coo = d3[g][c1][c2]
                for l1 in coo.mat:
                    for l2, v in coo.mat[l1].items():
                        out.append((int(g), int(c1), int(c2), int(l1), int(l2), float(v)))
    return np.array(out, np.float64).reshape(-1, 6)


def dump_history(epi, n_param_days=10):
    st = epi.history.states
    out = dict(
        X=np.stack([np.array(s.obs.current_comp) for s in st]),
        cost=np.stack([np.array(s.obs.cumulative_cost) for s in st]),
    )
    k = min(n_param_days, len(st))
    out["p_mp"] = np.stack([np.array(s.unobs.model_parameters) for s in st[:k]])
    out["p_ft"] = np.stack([np.array(s.unobs.facility_timespent) for s in st[:k]])
    out["p_fi"] = np.stack([np.array(s.unobs.facility_interaction) for s in st[:k]])
    out["p_cost"] = np.stack([np.array(s.unobs.delta_cost) for s in st[:k]])
    out["p_mob"] = np.stack([np.concatenate([np.array(m).reshape(-1) for m in s.unobs.mode_reduced_csr]) for s in st[:k]])
    for d in range(k):
        out[f"p_moves{d}"] = moves_of(st[d].unobs)
    # a late day too (post S-exhaustion regimes)
    return out


def run_env_plan(env, actions):
    """weekly actions through EpiEnv.step (epipolicy_environment.py:39-62)"""
    _trace.clear()
    env.reset()
    rewards, obs_w = [], []
    for a in actions:
        obs, r, done, _ = env.step(a)
        rewards.append(r)
        obs_w.append(obs)
        if done:
            break
    out = dump_history(env.epi)
    out["actions"] = np.asarray(actions, np.float32)
    out["rewards"] = np.array(rewards)
    out["obs_weekly"] = np.stack(obs_w)
    out["nfev"] = np.array([t[0] for t in _trace], np.int32)
    out["nsteps"] = np.array([t[1] for t in _trace], np.int32)
    return out


def run_schedule_plan(epi, T):
    """the scenario's own schedule through daily Epidemic.step (epidemic.py:125-128)"""
    _trace.clear()
    epi.reset()
    rewards, active, cpv = [], [], []
    st = epi.static
    off = np.cumsum([0] + [len(i.cp_list) for i in st.interventions])
    for t in range(T):
        action = st.schedule.action_list[t]
        act_row = np.zeros(st.intervention_count, np.int32)
        cpv_row = np.array([c.default_value for i in st.interventions for c in i.cp_list], np.float32)
        for act in action:
            act_row[act.index] = 1
            cpv_row[off[act.index]:off[act.index + 1]] = np.asarray(act.cpv_list)
        active.append(act_row)
        cpv.append(cpv_row)
        _, r, done = epi.step(action)
        rewards.append(r)
    out = dump_history(epi)
    out["rewards"] = np.array(rewards)
    out["plan_active"] = np.stack(active)
    out["plan_cpv"] = np.stack(cpv)
    out["nfev"] = np.array([t[0] for t in _trace], np.int32)
    out["nsteps"] = np.array([t[1] for t in _trace], np.int32)
    return out


def lax_action(env):
    # baselines.py:35-39 ("lax"): vaccination at a fixed rate, everything else fully on
    a = np.ones(env.action_space.shape[0], np.float32)
    a[0] = 2224526 * 0.9 / (12 * 30) / 9200
    return a


def main(names):
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        t0 = time.time()
        session = json.load(open(os.path.join(REF, "jsons", name + ".json")))
        has_opt = "optimize" in session
        env = EpiEnv(session)
        epi = env.epi
        d = export_static(epi)
        d.meta["scenario"] = name
        d.save(os.path.join(SCN, f"desc_{name}.npz"))
        out = {f"init_{k[5:]}": v for k, v in export_initial_parameter(epi).items()}
        A = env.action_space.shape[0]
        n_steps = (d.horizon + 6) // 7
        if has_opt:
            rng = np.random.default_rng(7)
            plans = {"rnd": rng.uniform(0, 1, (n_steps, A)).astype(np.float32),
                     "lax": np.tile(lax_action(env), (n_steps, 1))}
            for pname, acts in plans.items():
                for k, v in run_env_plan(env, acts).items():
                    out[f"{pname}_{k}"] = v
        for k, v in run_schedule_plan(epi, d.horizon if has_opt else 14).items():
            out[f"sched_{k}"] = v
        if name == "warmup_scenario":
            # the one recorded number in the reference repo: user.ipynb:82,93
            epi.reset()
            _, r, _ = epi.step([construct_act(0, [0.2, 10, 10])])
            out["known_answer_reward"] = np.array([r])
            print("warm-up known answer", repr(r), "(user.ipynb:82 says -4566.504778395923)")
        if name == "SIRV_B":
            # BASELINE.json configs[1]: env e uses default_rng(1000+e).uniform(0,1,(53,4)); a subset of envs
            sub = []
            for e in range(16):
                acts = np.random.default_rng(1000 + e).uniform(0, 1, (n_steps, A)).astype(np.float32)
                o = run_env_plan(env, acts)
                sub.append(o)
            out["batch_X"] = np.stack([o["X"] for o in sub])
            out["batch_cost"] = np.stack([o["cost"] for o in sub])
            out["batch_rewards"] = np.stack([o["rewards"] for o in sub])
            out["batch_nfev"] = np.stack([o["nfev"] for o in sub])
        np.savez_compressed(os.path.join(OUT, f"traj_{name}.npz"), **out)
        print(f"{name}: done in {time.time() - t0:.1f}s  n_flat={d.n_flat} ops={d.NOP}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["SIR_A", "SIR_B", "SIRV_A", "SIRV_B", "COVID_A", "COVID_B", "COVID_C", "warmup_scenario"])
