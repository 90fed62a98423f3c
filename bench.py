#!/usr/bin/env python
"""bench.py — env-steps/s of the batched epidemic step (BASELINE.json metric) on N B200s.

A "step" is one gym step (EpiEnv.step: 7 simulated days with one action, fused in ONE kernel launch)
over the rank's whole batch of environments. Weak scaling: every rank owns `--envs` environments of
the same scenario; at N > 1 the observations / rewards / done flags of every shard are delivered to the learner
rank each step: the step kernel's epilogue stores them into rank 0's receive buffer over NVLink (symmetric memory;
one device-side barrier per step), or, with --gather nccl / when symmetric memory is unavailable, one NCCL
all-gather per step as the north_star words it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload syn|covid_c|sirv_b|syn_s|syn_m] [--envs E]
  python bench.py --impl reference ...   # the CPU restatement of the reference on all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PERIOD = 7
WORKLOADS = {
    # name: (shipped scenario, default envs per GPU, description)
    "syn": ("syn", 4096, "synthetic COVID-like L=1000 G=16 F=8 x 4096 envs per GPU, fp32, random actions "
                         "(BASELINE.json configs[3]: 65536 envs = --envs 8192 at 8 GPUs; the 4096-env batch is the one "
                         "the north_star roofline target is quoted on)"),
    "covid_c": ("COVID_C", 4096, "COVID_C.json x 4096 envs per GPU, fp32, random actions (BASELINE.json configs[2])"),
    "sirv_b": ("SIRV_B", 1024, "SIRV_B.json x 1024 envs per GPU (BASELINE.json configs[1])"),
    "syn_s": ("syn_s", 4096, "synthetic COVID-like L=20 G=4 F=8 (epipolicy_rl_b200/synth.py)"),
    "syn_m": ("syn_m", 1024, "synthetic COVID-like L=100 G=8 F=8 (epipolicy_rl_b200/synth.py)"),
}


def load_desc(workload):
    """shipped scenario descriptor (epipolicy_rl_b200/scenarios); pure Python, loads no native library"""
    from epipolicy_rl_b200 import scenarios
    return scenarios.load(WORKLOADS[workload][0] if workload in WORKLOADS else workload)


def workload_config(args, desc, n_envs):
    """the `config` object of the JSON line: identical in both arms (it names the workload, nothing device-specific)"""
    d = WORKLOADS.get(args.workload)
    name = d[2] if d else f"scenario {args.workload}: L={desc.L} G={desc.G} F={desc.F} C={desc.C}"
    state_gb = n_envs * (desc.C * desc.L * desc.G * 8 * 2) / 1e9
    return {"workload": name, "scenario": args.workload, "envs_per_gpu": n_envs, "days_per_step": PERIOD,
            "precision": args.precision, "actions": "U(0,1)^A per env and gym step, seed 1234 + rank",
            "l2": f"per-step working set {state_gb:.2f} GB of env state (> 126 MB L2 when > 0.13) and a 256 MiB flush "
                  "write between timed GPU steps"}


def algorithmic_bytes(desc, info, w, fused_days):
    """bytes one env must move per launch of `fused_days` days (X, yesterday's move
    table, previous-state rows and actions once per launch; cost arrays once per launch too since they
    stay on chip between the fused days), plus the variant that also counts the cumulative_move
    accumulators the mirrored RK45 error norm needs (bytes_ctrl)."""
    C, L, G, I, A = desc.C, desc.L, desc.G, desc.I, desc.A
    LG = L * G
    n_move, n_chg = info.n_slot, 1 if any(desc.sel_arg.reshape(-1, 4)[:, 3] == 2) else 0
    base = w * (2 * C * LG + 2 * n_move * LG + 2 * LG * n_chg + 3 * A) + 8 * 4 * I * L + w + 1
    ctrl = base + w * 2 * info.n_acc * LG
    return base, ctrl


def algorithmic_flops(desc, info, nfev_per_day, fused_days):
    """fp32 flops one env needs per launch, after the obvious algebra (one force-of-infection
    numerator per distinct weighted infectious sum n_w, one shared denominator): per right-hand-side evaluation
    2 nnz G (n_w+1) [CSC gathers] + 2 L F G^2 (n_w+1) [G x G mixing] + 3 L G F n_w + 2 nnz G n_w [CSR gather]
    + 4 E L G + 2 C L G [edges, scatter, alive], plus the RK45 stage algebra 2*7*C*L*G per evaluation."""
    C, L, G, F = desc.C, desc.L, desc.G, desc.F
    nnz, E, nw = int(desc.scalars["nnz"]), int(desc.scalars["E"]), max(1, info.n_groups_inf)
    rhs = (2 * nnz * G * (nw + 1) + 2 * L * F * G * G * (nw + 1) + 3 * L * G * F * nw + 2 * nnz * G * nw
           + 4 * E * L * G + 2 * C * L * G)
    return fused_days * nfev_per_day * (rhs + 2 * 7 * C * L * G)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [s for t, s in self.samples if t0 <= t <= t1] or [s for _, s in self.samples[-3:]]
        sm, smax, reasons = [], None, set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(rows)}


def _sample_plan(one_env_step_s, budget_s, want_envs, want_steps):
    """bounded CPU sample: (n_envs, n_steps, days) that costs about `budget_s` seconds on one thread"""
    if one_env_step_s * want_envs * want_steps <= budget_s:
        return want_envs, want_steps, PERIOD
    if one_env_step_s <= budget_s:
        n_steps = max(1, min(want_steps, int(budget_s / one_env_step_s / max(want_envs, 1))))
        return max(1, min(want_envs, int(budget_s / one_env_step_s / n_steps))), n_steps, PERIOD
    # a whole gym step of one env does not fit the budget (SYN: about a minute): time single simulated days and
    # scale by the 7 days of a gym step
    return 1, 1, max(1, min(PERIOD, int(budget_s / (one_env_step_s / PERIOD))))


def cpu_port_rate(desc, n_envs, n_steps, seed=0, budget_s=12.0):
    """The oracle (C restatement of the reference), one host thread: env-steps/s on a bounded sample."""
    from oracle.oracle import rollout
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    rollout(desc, rng.uniform(0, 1, (1, 1, desc.A)).astype(np.float32), 1, want_final=False)
    one = (time.perf_counter() - t0) * PERIOD
    n_envs, n_steps, days = _sample_plan(one, budget_s, n_envs, n_steps)
    acts = rng.uniform(0, 1, (n_envs, n_steps, desc.A)).astype(np.float32)
    t0 = time.perf_counter()
    rollout(desc, acts, days, want_final=False)
    dt = time.perf_counter() - t0
    rate = n_envs * n_steps * (days / PERIOD) / dt
    sample = f"{n_envs} envs x {n_steps} gym steps" + ("" if days == PERIOD else f" x {days} of 7 days (scaled by days/7)")
    return rate, dt, sample


def _ref_worker(args):
    workload, n_envs, n_steps, days, seed = args
    desc = load_desc(workload)
    from oracle.oracle import rollout
    acts = np.random.default_rng(seed).uniform(0, 1, (n_envs, n_steps, desc.A)).astype(np.float32)
    rollout(desc, acts, days, want_final=False)
    return n_envs * n_steps * days / PERIOD


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on every host core, same workload / metric / config. The
    reference is Python + numba and cannot travel to the GPU box, so what runs is its C restatement
    (oracle/, cpu_baseline.kind = "port", about 56x faster per core than the Python reference). Each step is a bounded
    sample of the batch; `value` is normalised to env-steps/s. Loads no GPU library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    desc = load_desc(args.workload)
    n_envs = args.envs or (WORKLOADS[args.workload][1] if args.workload in WORKLOADS else 4096)
    cores = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    _ref_worker((args.workload, 1, 1, 1, 7))
    one = (time.perf_counter() - t0) * PERIOD
    per_core, _, days = _sample_plan(one, args.ref_seconds_per_step, args.ref_envs_per_core, 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        job = [(args.workload, per_core, 1, days, 100 + i) for i in range(cores)]
        # bound the whole run (the contract: a default --steps K --warmup W run ends within a few minutes): on SYN the
        # smallest possible step - one simulated day of one env per core - already takes ~25 s, so the number of timed
        # steps is capped by --ref-total-seconds and reported as `steps` (the request as `steps_requested`)
        step_est = 1.25 * (one / PERIOD) * days * per_core
        n_warm = max(min(args.warmup, 2), 1) if step_est < 5.0 else 1
        n_steps = min(args.steps, max(2, int((args.ref_total_seconds - n_warm * step_est) / step_est)))
        for _ in range(n_warm):
            pool.map(_ref_worker, job)
        t0 = time.perf_counter()
        done = 0.0
        steps_requested, args.steps, args.warmup = args.steps, n_steps, n_warm
        for k in range(args.steps):
            done += sum(pool.map(_ref_worker, [(args.workload, per_core, 1, days, 1000 + k * cores + i) for i in range(cores)]))
        dt = time.perf_counter() - t0
    v = done / dt
    sample = (f"{per_core * cores} of the {n_envs} envs per step ({per_core} per core), "
              + ("1 gym step (7 days)" if days == PERIOD else f"{days} of the 7 days of a gym step (scaled by days/7)")
              + f", {args.steps} steps, {cores} processes of oracle/epi_oracle.c")
    line = {"impl": "reference", "metric": "env-steps/s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, desc, n_envs), "sim_days_per_s": v * PERIOD,
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
                             "python_reference_survey": PY_REF},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "steps_requested": steps_requested}
    print(json.dumps(line), flush=True)


# the unmodified Python reference, timed in the build container (it cannot run on the GPU box)
PY_REF = {"covid_c_env_steps_per_s_per_core": 8.5, "syn_seconds_per_env_step": 194.6,
          "where": "SURVEY.md §6, Appendix F: build container, one core, after the 105 s numba JIT"}


def timed_steps(eng, torch, dist, world, acts, K, W, gather, flush):
    """W warm-up steps, then K steps timed with CUDA events on the launching stream (L2 flushed outside the event
    pairs), barrier + synchronize on both sides, max over ranks. Returns (ms_per_step, launches, t_wall0, t_wall1)."""
    def one_step(i):
        obs, rew, done = eng.step(acts[i], PERIOD)
        if gather is not None:
            gather(obs, rew, done)
            gather.release()

    for i in range(W):
        one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = eng.launches
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    resets = 0
    for i in range(K):
        flush.zero_()            # evict L2 between timed iterations (outside the event pair)
        ev[i][0].record()
        one_step(W + i)
        ev[i][1].record()
        if (W + i + 1) % 45 == 0:    # episodes are 53 steps long: restart before the horizon (untimed)
            torch.cuda.synchronize()
            eng.reset()
            resets += 1
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    return dev_ms / K, eng.launches - launches0 - resets, t_wall0, t_wall1


def verify_gather(eng, gather, torch, dist, world, rank):
    """after the timed region: what the learner rank holds in its gather buffer against one NCCL exchange of every
    rank's local results, bit for bit (obs | reward | done of the last step)"""
    got = gather.last() if rank == 0 else None
    ok = True
    for r in range(world):
        for idx, t in enumerate((eng.obs, eng.reward, eng.done)):
            tmp = t.clone() if rank == r else torch.empty_like(t)
            dist.broadcast(tmp, src=r)
            if rank == 0:
                ok = ok and bool(torch.equal(got[r][idx], tmp))
            del tmp
    return ok


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from epipolicy_rl_b200.engine import BatchedEpidemic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    desc = load_desc(args.workload)
    n_envs = args.envs or (WORKLOADS[args.workload][1] if args.workload in WORKLOADS else 4096)
    eng = BatchedEpidemic(desc, n_envs, device=local, precision=args.precision)
    info = eng.info
    w = 4 if args.precision == "fp32" else 8
    K, W = args.steps, max(args.warmup, 3)
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    acts_host = torch.rand((K + W, n_envs, desc.A), generator=gen)
    acts = acts_host.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    gather, gather_kind = None, "none"
    if world > 1:
        # results of every shard go to the learner rank: by default the step kernel's epilogue stores them straight
        # into rank 0's receive buffer over NVLink (dist.PeerGather); --gather nccl = one all-gather per gym step
        from epipolicy_rl_b200.dist import make_gather
        gather, gather_kind = make_gather(eng, world, prefer_peer=args.gather == "peer")

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    ms_per_step, launches, t_wall0, t_wall1 = timed_steps(eng, torch, dist, world, acts, K, W, gather, flush)
    clocks = sampler.stop(t_wall0, t_wall1)
    value = world * n_envs / (ms_per_step * 1e-3)
    nfev = float(eng.get_state("trace")[:, 0].mean())
    gather_ok = verify_gather(eng, gather, torch, dist, world, rank) if world > 1 else None
    if gather is not None and hasattr(gather, "close"):
        gather.close()

    # ---- end to end through the host-buffer C-ABI call (H2D actions, launch, D2H obs/reward/done)
    eng.reset()
    acts_np = acts_host.numpy()
    for i in range(2):
        eng.step_host(acts_np[i], PERIOD)
    Ke = min(K, 30)
    t0 = time.perf_counter()
    for i in range(Ke):
        eng.step_host(acts_np[W + i], PERIOD)
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e = world * n_envs * Ke / e2e_dt
    rs = 4 if args.precision == "fp32" else 8
    h2d, d2h = n_envs * desc.A * 4, n_envs * (info.obs_dim * rs + 8 + 1)
    kernel_name = {0: "warp-team", 1: "cta-team", 2: "cell", 3: "env", 4: "env-group"}.get(info.team_kind, "?") + \
        ("+specialised" if info.specialized else "")
    launch = {"kernel": kernel_name, "threads": info.threads, "grid": info.grid, "smem_bytes": info.smem_bytes,
              "collective": gather_kind}
    eng.close()
    del eng

    # ---- secondary workload (flat keys): COVID_C x 4096 envs, the shipped-scenario configuration (configs[2])
    sec = {}
    if rank == 0 and args.workload == "syn" and not args.no_secondary:
        d2 = load_desc("covid_c")
        e2 = BatchedEpidemic(d2, 4096, device=local, precision=args.precision)
        a2 = torch.rand((43, 4096, d2.A), generator=gen).cuda()
        ms2, _, _, _ = timed_steps(e2, torch, dist, 1, a2, 40, 3, None, flush)
        b2, c2 = algorithmic_bytes(d2, e2.info, w, PERIOD)
        sec = {"covid_c_env_steps_per_s": 4096 / (ms2 * 1e-3), "covid_c_ms_per_step": ms2, "covid_c_envs": 4096,
               "covid_c_bytes_per_env_launch": b2, "covid_c_hbm_frac": 4096 * b2 / (ms2 * 1e-3) / 1e9 / hbm_peak()[0]}
        e2.close()

    if rank == 0:
        base_b, ctrl_b = algorithmic_bytes(desc, info, w, PERIOD)
        peak, peak_src = hbm_peak()
        achieved = n_envs * base_b / (ms_per_step * 1e-3) / 1e9
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            ent = json.load(open(tpath)).get(f"{args.workload}:{args.precision}:{n_envs}", {})
            traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
        # the second roofline (SURVEY.md §8d asks for both): algorithmic fp32 flops against the CUDA-core peak
        # 148 SMs x 128 lanes x 2 x SM clock; nfev of the last simulated day of this run, averaged over the envs
        flops = algorithmic_flops(desc, info, nfev, PERIOD)
        sm_ghz = (clocks.get("sm_max_mhz") or 1965.0) / 1e3
        fp32_peak = 148 * 128 * 2 * sm_ghz / 1e3  # TFLOP/s
        fp32_ach = n_envs * flops / (ms_per_step * 1e-3) / 1e12
        cpu_v, cpu_dt, cpu_sample = cpu_port_rate(desc, args.cpu_sample_envs, args.cpu_sample_steps)
        dram_frac = traffic / (ms_per_step * 1e-3) / 1e9 / peak if traffic else None
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": workload_config(args, desc, n_envs), "launch": launch, "sim_days_per_s": value * PERIOD,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "bytes_per_env_launch": base_b, "bytes_per_env_launch_with_rk45_accumulators": ctrl_b,
                         "note": "algorithmic bytes = state in / state out once per 7-day launch (SURVEY.md §8d); the "
                                 "kernel keeps X in f64 in both modes (the fp32 mode's float copy is a second column), "
                                 "so it cannot reach this denominator on the X term; with ~92 right-hand-side evaluations "
                                 "per gym step the path is CUDA-core / latency bound, see fp32_flop_frac and dram_frac"},
            # flat copies (the driver's parser keeps top-level scalars)
            "bytes_day": base_b, "bytes_day_ctrl": ctrl_b, "hbm_frac_algorithmic": achieved / peak,
            "hbm_frac_ctrl": n_envs * ctrl_b / (ms_per_step * 1e-3) / 1e9 / peak,
            "dram_bytes_per_env_step": (traffic / n_envs) if traffic else None, "dram_frac": dram_frac,
            "fp32_flop_frac": fp32_ach / fp32_peak, "fp32_tflops": fp32_ach, "fp32_peak_tflops": fp32_peak,
            "flops_per_env_launch": flops, "nfev_per_day": nfev,
            "cpu_baseline": {"value": cpu_v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                             "sample": f"{cpu_sample}, {cpu_dt:.1f} s, oracle/epi_oracle.c single thread",
                             "python_reference_survey": PY_REF},
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if gather_ok is not None:
            line["gather_verified"] = bool(gather_ok)
        line.update(sec)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="syn")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="N > 1: how results reach the learner rank")
    ap.add_argument("--no-secondary", action="store_true", help="skip the COVID_C x 4096 secondary measurement")
    ap.add_argument("--cpu-sample-envs", type=int, default=64)
    ap.add_argument("--cpu-sample-steps", type=int, default=53)
    ap.add_argument("--ref-envs-per-core", type=int, default=64)
    ap.add_argument("--ref-total-seconds", type=float, default=240.0,
                    help="--impl reference: bound on the whole run; caps the number of timed steps when one step is slow")
    ap.add_argument("--ref-seconds-per-step", type=float, default=3.0,
                    help="--impl reference: CPU seconds per core and step the bounded sample aims at")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)   # CPU only: builds and loads oracle/ alone
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
