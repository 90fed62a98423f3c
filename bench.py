#!/usr/bin/env python
"""bench.py — env-steps/s of the batched epidemic step (BASELINE.json metric) on N B200s.

A "step" is one gym step (EpiEnv.step: 7 simulated days with one action, fused in ONE kernel launch)
over the rank's whole batch of environments. Weak scaling: every rank owns `--envs` environments of
the same scenario; at N > 1 the observations / rewards / done flags of every shard are delivered to the learner
rank each step: the step kernel's epilogue stores them into rank 0's receive buffer over NVLink (symmetric memory;
one device-side barrier per step), or, with --gather nccl / when symmetric memory is unavailable, one NCCL
all-gather per step as the north_star words it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload covid_c|syn_small|syn] [--envs E]
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
    # name: (desc file or generator, default envs per GPU, description)
    "covid_c": ("tests/golden/desc_COVID_C.npz", 4096, "COVID_C.json x 4096 envs per GPU, fp32, random actions (BASELINE.json configs[2])"),
    "sirv_b": ("tests/golden/desc_SIRV_B.npz", 1024, "SIRV_B.json x 1024 envs per GPU (BASELINE.json configs[1])"),
    "syn_s": ("tests/golden/desc_syn_s.npz", 4096, "synthetic COVID-like L=20 G=4 F=8 (epipolicy_rl_b200/synth.py)"),
    "syn_m": ("tests/golden/desc_syn_m.npz", 1024, "synthetic COVID-like L=100 G=8 F=8 (epipolicy_rl_b200/synth.py)"),
    "syn": ("tests/golden/desc_syn.npz", 296, "synthetic COVID-like L=1000 G=16 F=8 (BASELINE.json configs[3])"),
}


def load_desc(workload):
    from epipolicy_rl_b200.desc import EpiDesc
    if workload in WORKLOADS:
        return EpiDesc.load(os.path.join(ROOT, WORKLOADS[workload][0]))
    from epipolicy_rl_b200 import synth
    return synth.workload_desc(workload)


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


def cpu_port_rate(desc, n_envs, n_steps, seed=0, budget_s=12.0):
    """The oracle (C restatement of the reference), one host thread: env-steps/s on a bounded sample
    (at most n_envs x n_steps gym steps, cut down to about `budget_s` seconds of CPU work)."""
    from oracle.oracle import rollout
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    rollout(desc, rng.uniform(0, 1, (1, 1, desc.A)).astype(np.float32), PERIOD, want_final=False)
    one = time.perf_counter() - t0
    if one * n_envs * n_steps > budget_s:
        n_steps = max(1, min(n_steps, int(budget_s / one / max(n_envs, 1))))
        n_envs = max(1, min(n_envs, int(budget_s / one / n_steps)))
    acts = rng.uniform(0, 1, (n_envs, n_steps, desc.A)).astype(np.float32)
    t0 = time.perf_counter()
    rollout(desc, acts, PERIOD, want_final=False)
    dt = time.perf_counter() - t0
    return n_envs * n_steps / dt, dt, n_envs, n_steps


def _ref_worker(args):
    desc_path_or_name, n_envs, n_steps, seed = args
    desc = load_desc(desc_path_or_name)
    from oracle.oracle import rollout
    acts = np.random.default_rng(seed).uniform(0, 1, (n_envs, n_steps, desc.A)).astype(np.float32)
    rollout(desc, acts, PERIOD, want_final=False)
    return n_envs * n_steps


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference cannot travel to
    the GPU box) on every host core, same workload/metric; each step = a bounded sample of the batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    desc = load_desc(args.workload)
    cores = len(os.sched_getaffinity(0))
    per_core = args.ref_envs_per_core
    # bounded sample: about 2 s of CPU work per core and step
    t0 = time.perf_counter()
    _ref_worker((args.workload, 1, 1, 7))
    one = time.perf_counter() - t0
    per_core = max(1, min(per_core, int(2.0 / max(one, 1e-6))))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        job = [(args.workload, per_core, 1, 100 + i) for i in range(cores)]
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker, job)
        t0 = time.perf_counter()
        done = 0
        for k in range(args.steps):
            done += sum(pool.map(_ref_worker, [(args.workload, per_core, 1, 1000 + k * cores + i) for i in range(cores)]))
        dt = time.perf_counter() - t0
    v = done / dt
    sample = f"{per_core * cores} envs x 1 gym step (7 days) per step, {args.steps} steps, {cores} processes"
    line = {"impl": "reference", "metric": "env-steps/s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, desc), "sim_days_per_s": v * PERIOD},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args, desc):
    d = WORKLOADS.get(args.workload)
    base = d[2] if d else f"synthetic COVID-like {args.workload}: L={desc.L} G={desc.G} F={desc.F} C={desc.C}"
    return base


def run_ours(args):
    import torch
    import torch.distributed as dist
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
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    acts_host = torch.rand((K + W, n_envs, desc.A), generator=g)
    acts = acts_host.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    gather, gather_kind = None, "none"
    if world > 1:
        # results of every shard go to the learner rank: by default the step kernel's epilogue stores them straight
        # into rank 0's receive buffer over NVLink (dist.PeerGather); --gather nccl = one all-gather per gym step
        from epipolicy_rl_b200.dist import make_gather
        gather, gather_kind = make_gather(eng, world, prefer_peer=args.gather == "peer")

    def one_step(i):
        obs, rew, done = eng.step(acts[i], PERIOD)
        if gather is not None:
            gather(obs, rew, done)

    for i in range(W):
        one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = eng.launches
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()            # evict L2 between timed iterations (outside the event pair)
        ev[i][0].record()
        one_step(W + i)
        ev[i][1].record()
        if (i + 1) % 30 == 0:    # episodes are 53 steps long: restart before the horizon (untimed)
            torch.cuda.synchronize()
            eng.reset()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = eng.launches - launches0 - (K // 30)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / K
    value = world * n_envs / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (H2D actions, launch, D2H obs/reward/done)
    eng.reset()
    acts_np = acts_host.numpy()
    for i in range(3):
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

    if rank == 0:
        base_b, ctrl_b = algorithmic_bytes(desc, info, w, PERIOD)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = n_envs * base_b / (ms_per_step * 1e-3) / 1e9
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{args.workload}:{args.precision}:{n_envs}", {}).get("dram_bytes_per_launch")
        # the second roofline (SURVEY.md §8d asks for both): algorithmic fp32 flops against the CUDA-core peak
        # 148 SMs x 128 lanes x 2 x SM clock; nfev of the last simulated day of this run, averaged over the envs
        nfev = float(eng.get_state("trace")[:, 0].mean())
        flops = algorithmic_flops(desc, info, nfev, PERIOD)
        sm_ghz = (clocks.get("sm_max_mhz") or 1965.0) / 1e3
        fp32_peak = 148 * 128 * 2 * sm_ghz / 1e3  # TFLOP/s
        fp32_ach = n_envs * flops / (ms_per_step * 1e-3) / 1e12
        cpu_v, cpu_dt, cpu_ne, cpu_ns = cpu_port_rate(desc, args.cpu_sample_envs, args.cpu_sample_steps)
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, desc), "envs_per_gpu": n_envs, "days_per_step": PERIOD,
                       "sim_days_per_s": value * PERIOD, "l2": "flushed between timed steps (256 MiB write)",
                       "kernel": {0: "warp-team", 1: "cta-team", 2: "cell", 3: "env"}[info.team_kind] +
                                 ("+specialised" if info.specialized else ""),
                       "threads": info.threads, "grid": info.grid,
                       "smem_bytes": info.smem_bytes, "collective": gather_kind},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_env_launch": base_b,
                         "bytes_per_env_launch_with_rk45_accumulators": ctrl_b,
                         "frac_with_rk45_accumulators": n_envs * ctrl_b / (ms_per_step * 1e-3) / 1e9 / peak,
                         "fp32_cuda_core": {"achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s",
                                            "frac": fp32_ach / fp32_peak, "flops_per_env_launch": flops,
                                            "nfev_per_day": nfev}},
            "cpu_baseline": {"value": cpu_v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                             "sample": f"{cpu_ne} envs x {cpu_ns} gym steps, {cpu_dt:.1f} s, "
                                       "oracle/epi_oracle.c single thread"},
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="covid_c")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="N > 1: how results reach the learner rank")
    ap.add_argument("--cpu-sample-envs", type=int, default=64)
    ap.add_argument("--cpu-sample-steps", type=int, default=53)
    ap.add_argument("--ref-envs-per-core", type=int, default=64)
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
