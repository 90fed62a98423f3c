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


class PeerGather:
    """Gather to the learner rank WITHOUT a collective call: every rank maps the learner's receive buffer (symmetric
    memory over NVLink / NVSwitch, torch.distributed._symmetric_memory) and binds its own slice of it as the
    engine's mirror output (engine.bind_mirror): the step kernel's epilogue stores each env's obs / reward / done
    into the learner's memory as the env finishes, overlapped with the envs still being stepped. What remains
    per gym step is one device-side barrier over the signal pads, so that the learner reads complete records.
    Same record layout as LearnerGather; `gathered()` is valid on the learner rank."""

    def __init__(self, eng, world: int, learner: int = 0, group=None):
        import torch.distributed._symmetric_memory as symm
        self.layout = PackedLayout(eng.n_envs, eng.obs_dim, eng.dtype)
        self.world, self.learner, self.rank = world, learner, dist.get_rank(group)
        n = self.layout.nbytes
        self.recv = symm.empty(world * n, dtype=torch.uint8, device=eng.device)
        self.recv.zero_()
        self.hdl = symm.rendezvous(self.recv, group if group is not None else dist.group.WORLD)
        peer = self.hdl.get_buffer(learner, (world * n,), torch.uint8)  # the learner's buffer, mapped here
        self.mine = self.layout.views(peer[self.rank * n:(self.rank + 1) * n])
        eng.bind_mirror(*self.mine)
        self.eng = eng

    def __call__(self, obs=None, rew=None, done=None):
        # the kernel has already stored this rank's records into the learner's buffer: order the reader behind every
        # rank's step kernel (stream-ordered barrier on the signal pads)
        self.hdl.barrier(channel=0)
        return self.gathered() if self.rank == self.learner else None

    def gathered(self):
        n = self.layout.nbytes
        return [self.layout.views(self.recv[r * n:(r + 1) * n]) for r in range(self.world)]

    def close(self):
        self.eng.bind_mirror(None, None, None)


def make_gather(eng, world: int, prefer_peer: bool = True):
    """PeerGather when symmetric memory can be set up on this box, else the NCCL all-gather. All ranks take the same
    branch (the outcome is agreed with an all-reduce)."""
    g, ok = None, 0
    if prefer_peer and eng.device.type == "cuda":
        try:
            g = PeerGather(eng, world)
            ok = 1
        except Exception as e:  # noqa: BLE001 - no symmetric memory on this box / build
            import sys
            print(f"[epi-b200] peer-memory gather unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
    t = torch.tensor([ok], device=eng.device, dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) == 1:
        return g, "peer-memory stores in the step kernel epilogue + signal-pad barrier (no collective call)"
    if g is not None:
        g.close()
    return LearnerGather(eng, world), "all_gather obs/reward/done"
