"""Multi-GPU plumbing: independent environments are sharded over ranks (one process per GPU) with no
collective inside the step; ONE NCCL all-gather per gym step delivers observations, rewards and done
flags of every shard to the learner. The engine writes its outputs straight into the
rank's slice of the packed send buffer, so the gather needs no staging copy."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of the env axis owned by `rank`: [lo, hi). Remainders go to the first ranks."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PackedLayout:
    """Byte layout of one rank's (obs | reward | done) record block, 16-byte aligned segments."""

    def __init__(self, n_envs: int, obs_dim: int, obs_dtype: torch.dtype):
        self.n_envs, self.obs_dim, self.obs_dtype = n_envs, obs_dim, obs_dtype
        es = torch.empty((), dtype=obs_dtype).element_size()
        a16 = lambda x: (x + 15) // 16 * 16  # noqa: E731
        self.o_obs, self.n_obs = 0, n_envs * obs_dim * es
        self.o_rew, self.n_rew = a16(self.n_obs), n_envs * 8
        self.o_done, self.n_done = a16(self.o_rew + self.n_rew), n_envs
        self.nbytes = a16(self.o_done + self.n_done)

    def views(self, buf: torch.Tensor):
        """(obs [N, obs_dim], reward [N] f64, done [N] u8) views of a uint8 buffer of `nbytes`."""
        obs = buf[self.o_obs:self.o_obs + self.n_obs].view(self.obs_dtype).view(self.n_envs, self.obs_dim)
        rew = buf[self.o_rew:self.o_rew + self.n_rew].view(torch.float64)
        done = buf[self.o_done:self.o_done + self.n_done]
        return obs, rew, done


class LearnerGather:
    """all-gather of the packed per-rank outputs. Works on NCCL (CUDA tensors) and gloo (CPU tensors, tests)."""

    def __init__(self, eng_or_layout, world: int, device=None, group=None):
        if isinstance(eng_or_layout, PackedLayout):
            self.layout, eng = eng_or_layout, None
        else:
            eng = eng_or_layout
            self.layout = PackedLayout(eng.n_envs, eng.obs_dim, eng.dtype)
            device = eng.device
        self.world, self.group = world, group
        self.send = torch.zeros(self.layout.nbytes, dtype=torch.uint8, device=device)
        self.recv = torch.zeros(world * self.layout.nbytes, dtype=torch.uint8, device=device)
        self.local = self.layout.views(self.send)
        if eng is not None:
            eng.bind_outputs(*self.local)  # the kernel epilogue writes into the send buffer

    def __call__(self, obs=None, rew=None, done=None):
        if obs is not None and obs.data_ptr() != self.local[0].data_ptr():
            self.local[0].copy_(obs); self.local[1].copy_(rew); self.local[2].copy_(done)
        if self.send.is_cuda:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            dist.all_gather(list(self.recv.chunk(self.world)), self.send, group=self.group)
        return self.gathered()

    def gathered(self):
        """[(obs, reward, done) of rank r for r in range(world)] as views of the receive buffer."""
        n = self.layout.nbytes
        return [self.layout.views(self.recv[r * n:(r + 1) * n]) for r in range(self.world)]
