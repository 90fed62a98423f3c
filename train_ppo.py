#!/usr/bin/env python
"""End-to-end PPO / SAC on GPU-batched epidemic environments (BASELINE.json configs[4]: COVID_A, envs sharded over the
GPUs of one box, one NCCL all-gather per gym step to the learner).

  python train_ppo.py [--scenario COVID_A] [--envs 4096] [--updates 20] [--algo ppo|sac]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 train_ppo.py ...

Prints one JSON line per update (rank 0) and a final summary with env-steps/s and the mean step reward of the
trained policy against the random and "lax" baselines of the reference's baselines.py:31-50.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="COVID_A")
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--updates", type=int, default=20)
    ap.add_argument("--n-steps", type=int, default=16)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--algo", default="ppo", choices=["ppo", "sac"], help="runner.py --algo")
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    import torch
    import torch.distributed as dist
    from epipolicy_rl_b200.desc import EpiDesc
    from epipolicy_rl_b200.env import BatchedEpiEnv
    from epipolicy_rl_b200.learner import BatchedPPO, BatchedSAC

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from epipolicy_rl_b200 import scenarios
    desc = scenarios.load(args.scenario)
    env = BatchedEpiEnv(desc, args.envs, device=local, precision=args.precision)
    pop = np.broadcast_to(desc.pop.reshape(1, desc.L, desc.G), (desc.C, desc.L, desc.G)).reshape(-1)
    dev = torch.device(f"cuda:{local}")

    def rollout_mean(policy, steps=53):
        """mean gym-step reward of a fixed policy over one episode of every env"""
        obs = env.reset()
        tot = torch.zeros(args.envs, dtype=torch.float64, device=dev)
        for _ in range(steps):
            obs, r, done, _ = env.step(policy(obs))
            tot += r
        return float(tot.mean()) / steps

    lax = torch.ones((args.envs, desc.A), device=dev)
    lax[:, 0] = 2224526 * 0.9 / (12 * 30) / 9200  # baselines.py:35-39
    base = {"random": rollout_mean(lambda o: torch.rand((args.envs, desc.A), device=dev)),
            "lax": rollout_mean(lambda o: lax)}
    log = lambda s: print(json.dumps(s), flush=True)  # noqa: E731
    if args.algo == "ppo":
        ppo = BatchedPPO(env, env.epi.obs_dim, desc.A, obs_scale=pop, n_steps=args.n_steps, lr=args.lr, device=dev)
        hist = ppo.train(1, log=log)  # first update untimed: cuBLAS / autograd / Adam state start-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist += ppo.train(args.updates, log=log)
        policy = lambda o: ppo.ac.act(ppo._norm(o), deterministic=True)[0].clamp(0, 1).contiguous()  # noqa: E731
    else:
        # the same number of gym steps as the PPO run: `updates` x `n_steps` collect steps, gradient steps after each
        sac = BatchedSAC(env, env.epi.obs_dim, desc.A, obs_scale=pop, lr=args.lr, batch_size=4096, gamma=0.98, tau=0.08,
                         gradient_steps=8, device=dev)
        hist = sac.train(args.n_steps)  # untimed start-up, as above
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist += sac.train(args.updates * args.n_steps, log=lambda s: log(s) if s["step"] % args.n_steps == 0 else None)
        policy = lambda o: sac.actor(sac._norm(o), deterministic=True, with_logprob=False)[0].contiguous()  # noqa: E731
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        trained = rollout_mean(policy)
    if rank == 0:
        print(json.dumps({"summary": True, "algo": args.algo, "scenario": args.scenario, "n_gpus": world, "envs_per_gpu": args.envs,
                          "updates": args.updates, "env_steps_per_s": world * args.envs * args.n_steps * args.updates / dt,
                          "train_seconds": dt, "mean_step_reward": {"trained": trained, **base},
                          "first_update_reward": hist[0]["mean_step_reward"], "last_update_reward": hist[-1]["mean_step_reward"]}),
              flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
