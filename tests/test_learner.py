"""Batched PPO learner (epipolicy_rl_b200/learner.py) on CPU: GAE against a plain loop, learning on a toy batched
env, and the 2-rank gloo path (all-gather of the packed records, learner update on rank 0, weight broadcast)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from epipolicy_rl_b200.learner import BatchedPPO, gae


class ToyEnv:
    """reward = -|a - target(obs)|^2 per env; episodes of 8 steps"""

    def __init__(self, n, seed=0):
        self.n, self.g = n, torch.Generator().manual_seed(seed)
        self.epi = self
        self.t = torch.zeros(n)
        self.obs = torch.rand((n, 3), generator=self.g)

    def reset(self, mask=None):
        m = torch.ones(self.n, dtype=torch.bool) if mask is None else mask.to(torch.bool)
        self.obs[m] = torch.rand((int(m.sum()), 3), generator=self.g)
        self.t[m] = 0
        return self.obs

    def step(self, a):
        target = self.obs[:, :2] * 0.5 + 0.25
        r = -((a - target) ** 2).sum(-1).double()
        self.t += 1
        self.obs = torch.rand((self.n, 3), generator=self.g)
        return self.obs, r, (self.t >= 8).to(torch.uint8), {}


def test_gae_matches_plain_loop():
    T, N = 5, 3
    g = torch.Generator().manual_seed(1)
    rew, val, last = torch.rand((T, N), generator=g), torch.rand((T, N), generator=g), torch.rand(N, generator=g)
    done = (torch.rand((T, N), generator=g) < 0.3).float()
    adv, ret = gae(rew, val, done, last, 0.9, 0.8)
    for n in range(N):
        run, nxt = 0.0, float(last[n])
        for t in reversed(range(T)):
            nt = 1.0 - float(done[t, n])
            delta = float(rew[t, n]) + 0.9 * nxt * nt - float(val[t, n])
            run = delta + 0.9 * 0.8 * nt * run
            nxt = float(val[t, n])
            assert abs(float(adv[t, n]) - run) < 1e-5
    assert torch.allclose(ret, adv + val)


def test_ppo_learns_on_toy_env():
    env = ToyEnv(256)
    ppo = BatchedPPO(env, 3, 2, obs_scale=np.ones(3), n_steps=8, n_epochs=4, lr=3e-3, reward_scale=1.0, device="cpu", seed=3)
    hist = ppo.train(25)
    assert hist[-1]["mean_step_reward"] > hist[0]["mean_step_reward"] + 0.02


class DetEnv:
    """deterministic batched env whose envs are told apart by their GLOBAL index only: env e of a fleet behaves the same
    whichever rank owns it (the property the GPU engine's identical-scenario envs have)"""

    def __init__(self, n, first, total):
        self.n, self.epi = n, self
        self.idx = torch.arange(first, first + n, dtype=torch.float32)
        self.total = total
        self.reset()

    def _obs(self):
        return torch.stack([self.s, self.s ** 2, self.idx / self.total], 1)

    def reset(self, mask=None):
        m = torch.ones(self.n, dtype=torch.bool) if mask is None else mask.to(torch.bool)
        if mask is None:
            self.s, self.t = (self.idx + 1) / (2 * self.total), torch.zeros(self.n)
        else:
            self.s = torch.where(m, (self.idx + 1) / (2 * self.total), self.s)
            self.t = torch.where(m, torch.zeros(self.n), self.t)
        self.obs = self._obs()
        return self.obs

    def step(self, a):
        r = -((a - self.s[:, None]) ** 2).sum(-1).double()
        self.s = 0.9 * self.s + 0.1 * a.mean(-1) + 0.01
        self.t = self.t + 1
        self.obs = self._obs()
        return self.obs, r, (self.t >= 5).to(torch.uint8), {}


def _det_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6
    env = DetEnv(n, rank * n, world * n)
    ppo = BatchedPPO(env, 3, 2, obs_scale=np.ones(3), n_steps=7, n_epochs=1, lr=1e-3, reward_scale=1.0, device="cpu", seed=5)
    batches = []
    for _ in range(2):  # two updates: the noise generator and the broadcast weights stay in step
        batch = ppo.collect()
        batches.append([x.clone() for x in batch])
        ppo.update(batch)
    if rank == 0:
        torch.save(batches, os.path.join(out, "multi.pt"))
    dist.destroy_process_group()


def test_two_rank_rollout_equals_one_rank_rollout(tmp_path):
    """identical envs on every rank (as the GPU engine's are): the W-rank fleet collects exactly what one rank over
    all W*N envs collects - per-env noise is a function of the global env index - and no two shards coincide"""
    port = 29450 + os.getpid() % 100
    mp.spawn(_det_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    multi = torch.load(tmp_path / "multi.pt")
    env = DetEnv(12, 0, 12)
    ppo = BatchedPPO(env, 3, 2, obs_scale=np.ones(3), n_steps=7, n_epochs=1, lr=1e-3, reward_scale=1.0, device="cpu", seed=5)
    for u in range(2):
        batch = ppo.collect()
        for a, b in zip(batch, multi[u]):
            assert a.shape == b.shape
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), u
        acts = multi[u][1]
        assert not torch.allclose(acts[:, :6], acts[:, 6:])  # the two shards explore differently
        ppo.update(batch)


def _sharded_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 16
    env = DetEnv(n, rank * n, world * n)
    ppo = BatchedPPO(env, 3, 2, obs_scale=np.ones(3), n_steps=5, n_epochs=2, lr=3e-3, reward_scale=1.0, device="cpu", seed=5,
                     update_mode="sharded")
    w0 = torch.cat([p.data.reshape(-1) for p in ppo.ac.parameters()]).clone()
    hist = ppo.train(6)
    flat = torch.cat([p.data.reshape(-1) for p in ppo.ac.parameters()])
    assert not torch.equal(flat, w0)
    torch.save((flat, [h["mean_step_reward"] for h in hist]), os.path.join(out, f"s{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_update_keeps_replicas_identical(tmp_path):
    """`sharded` update mode: no rollout exchange; gradients are averaged per minibatch step, so every rank applies the
    same optimiser steps and the replicas stay bit-identical without a weight broadcast; the logged reward is the
    fleet's."""
    port = 29350 + os.getpid() % 100
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    (w0, r0), (w1, r1) = torch.load(tmp_path / "s0.pt"), torch.load(tmp_path / "s1.pt")
    assert torch.equal(w0, w1)
    assert r0 == r1


def test_gae_matches_naive_fp64_loop_with_episode_ends():
    T, N = 16, 7
    g = torch.Generator().manual_seed(4)
    rew, val, last = torch.randn((T, N), generator=g), torch.randn((T, N), generator=g), torch.randn(N, generator=g)
    done = (torch.rand((T, N), generator=g) < 0.2).float()
    done[-1, 0] = 1.0
    adv, ret = gae(rew, val, done, last, 0.999, 0.99)
    r64, v64, d64, l64 = rew.double(), val.double(), done.double(), last.double()
    ref = torch.zeros((T, N), dtype=torch.float64)
    for n in range(N):
        for t in range(T):
            # naive definition: sum over l of (gamma*lambda)^l * delta_{t+l}, cut at the episode end
            acc, w = 0.0, 1.0
            for l in range(t, T):
                nxt = l64[n] if l == T - 1 else v64[l + 1, n]
                nt = 1.0 - d64[l, n]
                acc += w * (r64[l, n] + 0.999 * nxt * nt - v64[l, n])
                if d64[l, n] > 0:
                    break
                w *= 0.999 * 0.99
            ref[t, n] = acc
    assert torch.allclose(adv.double(), ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(ret, adv + val)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = ToyEnv(32, seed=rank)
    ppo = BatchedPPO(env, 3, 2, obs_scale=np.ones(3), n_steps=4, n_epochs=1, lr=1e-3, reward_scale=1.0, device="cpu", seed=5)
    batch = ppo.collect()
    assert batch[0].shape == (4, 64, 3)  # both shards, on every rank
    ppo.update(batch)
    flat = torch.cat([p.data.reshape(-1) for p in ppo.ac.parameters()])
    torch.save(flat, os.path.join(out, f"w{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gather_and_broadcast(tmp_path):
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w0, w1 = torch.load(tmp_path / "w0.pt"), torch.load(tmp_path / "w1.pt")
    assert torch.equal(w0, w1)


def test_squashed_gaussian_logprob_matches_torch_transform():
    from epipolicy_rl_b200.learner import SquashedGaussianActor
    torch.manual_seed(0)
    actor = SquashedGaussianActor(3, 2)
    obs = torch.rand(16, 3)
    a, logp = actor(obs)
    assert a.min() >= 0 and a.max() <= 1
    mu, log_std = actor.net(obs).chunk(2, -1)
    base = torch.distributions.Normal(mu, log_std.clamp(-20, 2).exp())
    tr = torch.distributions.TransformedDistribution(base, [torch.distributions.transforms.TanhTransform(),
                                                           torch.distributions.transforms.AffineTransform(0.5, 0.5)])
    ref = tr.log_prob(a.clamp(1e-6, 1 - 1e-6)).sum(-1)
    assert torch.allclose(logp, ref, atol=1e-3, rtol=1e-3)


def test_replay_ring_wraps():
    from epipolicy_rl_b200.learner import ReplayRing
    ring = ReplayRing(1, 1, 8, "cpu")
    for k in range(3):
        v = torch.arange(3, dtype=torch.float32)[:, None] + 10 * k
        ring.store(v, v, v[:, 0], v, torch.zeros(3))
    assert ring.n == 8 and ring.ptr == 1
    assert ring.o[0, 0] == 22 and ring.o[1, 0] == 1  # the 9th transition overwrote slot 0


def test_sac_learns_on_toy_env():
    from epipolicy_rl_b200.learner import BatchedSAC
    env = ToyEnv(128)
    sac = BatchedSAC(env, 3, 2, obs_scale=np.ones(3), gamma=0.0, lr=3e-3, batch_size=256, buffer_size=1 << 14,
                     gradient_steps=4, reward_scale=1.0, device="cpu", seed=3)
    hist = sac.train(120)
    first = np.mean([h["mean_step_reward"] for h in hist[:10]])
    last = np.mean([h["mean_step_reward"] for h in hist[-10:]])
    assert last > first + 0.02, (first, last)


def _sac_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from epipolicy_rl_b200.learner import BatchedSAC
    env = ToyEnv(32, seed=rank)
    sac = BatchedSAC(env, 3, 2, obs_scale=np.ones(3), gamma=0.0, lr=1e-3, batch_size=64, buffer_size=4096,
                     gradient_steps=2, reward_scale=1.0, device="cpu", seed=5 + rank)
    for _ in range(12):  # episodes of 8 steps: the masked restart path is exercised
        sac.collect_step()
        sac.update()
    if rank == 0:
        assert sac.ring.n == 12 * 64  # the transitions of both shards land in the learner's ring
    else:
        assert sac.ring is None
    flat = torch.cat([p.data.reshape(-1) for p in sac.actor.parameters()])
    torch.save(flat, os.path.join(out, f"a{rank}.pt"))
    dist.destroy_process_group()


def test_sac_two_rank_ring_on_learner_and_actor_broadcast(tmp_path):
    port = 29800 + os.getpid() % 150
    mp.spawn(_sac_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a0, a1 = torch.load(tmp_path / "a0.pt"), torch.load(tmp_path / "a1.pt")
    assert torch.equal(a0, a1)  # different seeds per rank, same actor after the broadcast
