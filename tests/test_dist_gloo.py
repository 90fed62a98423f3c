"""The N>1 host path on CPU: env sharding and the packed learner all-gather over gloo, world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n_total, obs_dim, out):
    sys.path.insert(0, ROOT)
    from epipolicy_rl_b200.dist import LearnerGather, PackedLayout, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    n = hi - lo
    # equal shard sizes are required by the single all-gather; pad the layout to the largest shard
    n_max = shard_range(n_total, 0, world)[1]
    g = LearnerGather(PackedLayout(n_max, obs_dim, torch.float32), world, device="cpu")
    obs, rew, done = g.local
    env_ids = torch.arange(lo, hi)
    obs[:n] = env_ids[:, None].float() * 10 + torch.arange(obs_dim)[None].float()
    rew[:n] = -env_ids.double()
    done[:n] = (env_ids % 2).to(torch.uint8)
    parts = g()
    if rank == 0:
        all_obs = torch.cat([parts[r][0][: shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0]] for r in range(world)])
        all_rew = torch.cat([parts[r][1][: shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0]] for r in range(world)])
        all_done = torch.cat([parts[r][2][: shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0]] for r in range(world)])
        np.savez(out, obs=all_obs.numpy(), rew=all_rew.numpy(), done=all_done.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from epipolicy_rl_b200.dist import shard_range
    for n, w in [(65536, 8), (4096, 1), (10, 4), (7, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


@pytest.mark.timeout(120)
def test_learner_gather_world2_gloo(tmp_path):
    out = str(tmp_path / "g.npz")
    n_total, obs_dim, world = 11, 5, 2
    mp.spawn(_worker, args=(world, 29500 + os.getpid() % 2000, n_total, obs_dim, out), nprocs=world, join=True)
    z = np.load(out)
    ids = np.arange(n_total)
    assert np.array_equal(z["obs"], ids[:, None] * 10.0 + np.arange(obs_dim)[None])
    assert np.array_equal(z["rew"], -ids.astype(np.float64))
    assert np.array_equal(z["done"], (ids % 2).astype(np.uint8))


class _FakeEngine:
    """what dist.make_gather / LearnerGather use of a BatchedEpidemic"""

    def __init__(self, n, obs_dim):
        self.n_envs, self.obs_dim, self.dtype, self.device = n, obs_dim, torch.float32, torch.device("cpu")
        self.obs = torch.zeros((n, obs_dim))
        self.reward = torch.zeros(n, dtype=torch.float64)
        self.done = torch.zeros(n, dtype=torch.uint8)

    def bind_outputs(self, obs, reward, done):
        self.obs, self.reward, self.done = obs, reward, done


def _fallback_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from epipolicy_rl_b200.dist import LearnerGather, make_gather
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(4, 3)
    g, kind = make_gather(eng, world)  # no CUDA device: every rank agrees on the all-gather
    assert isinstance(g, LearnerGather) and kind.startswith("all_gather")
    eng.obs.fill_(float(rank + 1))  # the "kernel" writes into the bound send buffer
    parts = g(eng.obs, eng.reward, eng.done)
    assert all(float(parts[r][0].mean()) == r + 1 for r in range(world))
    if rank == 0:
        open(out, "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_make_gather_falls_back_to_all_gather_without_peer_memory(tmp_path):
    out = str(tmp_path / "ok")
    mp.spawn(_fallback_worker, args=(2, 29300 + os.getpid() % 150, out), nprocs=2, join=True)
    assert os.path.exists(out)
