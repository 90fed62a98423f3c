#!/usr/bin/env python
"""End-to-end PPO / SAC on GPU-batched epidemic environments (BASELINE.json configs[4]: COVID_A, envs sharded over the
GPUs of one box, one NCCL all-gather per update delivers every shard's rollout to the learner rank).

  python train_ppo.py [--scenario COVID_A] [--envs 4096] [--updates 20] [--algo ppo|sac] [--jsonl profiles/ppo.jsonl]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 train_ppo.py ...

Prints one JSON line per update (rank 0; also appended to --jsonl) and a final summary with env-steps/s including the
learner, the env-only rate of the same fleet, and the mean step reward of the trained policy against the random, "lax"
and "aggressive" baselines of the reference's baselines.py:31-50.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_POP, MAX_DOSES = 2224526, 9200  # baselines.py:12-13


def vaccine_rate(total_pop, pc_pop, nmonths, max_doses):
    return total_pop * pc_pop / (nmonths * 30) / max_doses  # baselines.py:15-16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="COVID_A")
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--updates", type=int, default=20)
    ap.add_argument("--n-steps", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=2, help="PPO epochs per update (configs/ppo_default.yaml has n_epochs 1 for batches of 128 samples)")
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--algo", default="ppo", choices=["ppo", "sac"], help="runner.py --algo")
    ap.add_argument("--update", default="learner", choices=["learner", "sharded"],
                    help="learner: all-gather the rollouts, optimise on rank 0, broadcast; sharded: every rank optimises on its shard, gradients all-reduced")
    ap.add_argument("--no-graph", action="store_true", help="eager rollout instead of the captured CUDA graph")
    ap.add_argument("--jsonl", default="", help="append the per-update JSON lines and the summary to this file")
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    import torch
    import torch.distributed as dist
    from epipolicy_rl_b200 import scenarios
    from epipolicy_rl_b200.env import BatchedEpiEnv
    from epipolicy_rl_b200.learner import BatchedPPO, BatchedSAC

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    desc = scenarios.load(args.scenario)
    env = BatchedEpiEnv(desc, args.envs, device=local, precision=args.precision)
    pop = np.broadcast_to(desc.pop.reshape(1, desc.L, desc.G), (desc.C, desc.L, desc.G)).reshape(-1)
    dev = torch.device(f"cuda:{local}")
    logf = open(args.jsonl, "a") if (args.jsonl and rank == 0) else None

    def emit(rec):
        if rank == 0:
            line = json.dumps(rec)
            print(line, flush=True)
            if logf:
                logf.write(line + "\n"); logf.flush()

    def fleet_mean(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item()) / world

    def rollout_mean(policy, steps=53):
        """mean gym-step reward of a fixed policy over one episode of every env of the fleet; policy(obs, step)"""
        obs = env.reset()
        tot = torch.zeros(args.envs, dtype=torch.float64, device=dev)
        for i in range(steps):
            obs, r, done, _ = env.step(policy(obs, i))
            tot += r
        return fleet_mean(tot.mean()) / steps

    A = desc.A
    lax = torch.zeros((args.envs, A), device=dev)
    lax[:, 0] = vaccine_rate(TOTAL_POP, 0.9, 12, MAX_DOSES)  # baselines.py:35-39
    if A > 1:
        lax[:, 1] = 1

    def agg(obs, i):  # baselines.py:41-50, the step index compared as the reference does
        a = torch.zeros((args.envs, A), device=dev)
        if i <= 9 * 30:
            a[:, 0] = vaccine_rate(TOTAL_POP, 0.85, 9, MAX_DOSES)
        if A > 1:
            a[:, 1] = 0.8
        if A > 2:
            a[:, 2:] = 1.0 if i <= 4 * 30 else 0.5
        return a

    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    base = {"random": rollout_mean(lambda o, i: torch.rand((args.envs, A), device=dev, generator=gen)),
            "lax": rollout_mean(lambda o, i: lax), "aggressive": rollout_mean(agg)}
    # the same fleet without a learner: gym steps with random actions (the simulator's own rate)
    env.reset()
    acts = torch.rand((32, args.envs, A), device=dev, generator=gen)
    for i in range(4):
        env.step(acts[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(4, 32):
        env.step(acts[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    env_only = world * args.envs * 28 / (time.perf_counter() - t0)

    log = lambda s: emit({**s, "n_gpus": world})  # noqa: E731
    if args.algo == "ppo":
        ppo = BatchedPPO(env, env.epi.obs_dim, desc.A, obs_scale=pop, n_steps=args.n_steps, n_epochs=args.epochs,
                         minibatches=args.minibatches, lr=args.lr, device=dev, use_graph=not args.no_graph,
                         desync=(desc.horizon + 6) // 7, update_mode=args.update)
        hist = ppo.train(2, log=log)  # untimed: eager warm-up rollout, graph capture, cuBLAS / Adam state start-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist += ppo.train(args.updates, log=log)
        policy = lambda o, i: ppo.ac.act(ppo._norm(o), deterministic=True)[0].clamp(0, 1).contiguous()  # noqa: E731
        extra = {"rollout_graph": ppo._graph is not None, "update_graph": ppo._ugraph is not None}
    else:
        # the same number of gym steps as the PPO run: `updates` x `n_steps` collect steps, gradient steps after each
        sac = BatchedSAC(env, env.epi.obs_dim, desc.A, obs_scale=pop, lr=args.lr, batch_size=4096, gamma=0.98, tau=0.08,
                         gradient_steps=8, device=dev)
        hist = sac.train(args.n_steps)  # untimed start-up, as above
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist += sac.train(args.updates * args.n_steps, log=lambda s: log(s) if s["step"] % args.n_steps == 0 else None)
        policy = lambda o, i: sac.actor(sac._norm(o), deterministic=True, with_logprob=False)[0].contiguous()  # noqa: E731
        extra = {}
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        trained = rollout_mean(policy)
    rate = world * args.envs * args.n_steps * args.updates / dt
    emit({"summary": True, "algo": args.algo, "scenario": args.scenario, "n_gpus": world, "envs_per_gpu": args.envs,
          "updates": args.updates, "n_steps": args.n_steps, "env_steps_per_s": rate, "env_only_steps_per_s": env_only,
          "fraction_of_env_only": rate / env_only, "train_seconds": dt,
          "mean_step_reward": {"trained": trained, **base},
          "beats_lax": trained > base["lax"], "beats_random": trained > base["random"],
          "update_mode": args.update if args.algo == "ppo" else None, "epochs": args.epochs, "minibatches": args.minibatches,
          "first_update_reward": hist[0]["mean_step_reward"], "last_update_reward": hist[-1]["mean_step_reward"], **extra})
    # teardown: the captured graphs hold NCCL work (sharded mode all-reduces inside the update graph) and must go before
    # the process group does; a watchdog ends the process if the teardown still blocks (observed once at 8 GPUs: the
    # results above are complete and flushed at this point)
    import gc
    import threading
    sys.stdout.flush()
    if logf:
        logf.close()
    threading.Thread(target=lambda: (time.sleep(30), os._exit(0)), daemon=True).start()
    if args.algo == "ppo":
        ppo._graph = ppo._ugraph = None
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    env.close()


if __name__ == "__main__":
    main()
