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
        """[(obs, reward, done) of rank r for r in range(world)] as views of the receive buffer; valid until the next
        call (the next all-gather overwrites them, stream-ordered)."""
        n = self.layout.nbytes
        return [self.layout.views(self.recv[r * n:(r + 1) * n]) for r in range(self.world)]

    last = gathered

    def release(self):
        """nothing to release: the next all-gather is stream-ordered behind the readers"""


class PeerGather:
    """Gather to the learner rank WITHOUT a collective call: every rank maps the learner's receive buffer (symmetric
    memory over NVLink / NVSwitch, torch.distributed._symmetric_memory) and binds its own slice of it as the
    engine's mirror output (engine.bind_mirror): the step kernel's epilogue stores each env's obs / reward / done
    into the learner's memory as the env finishes, overlapped with the envs still being stepped. What remains
    per gym step is one device-side barrier over the signal pads, so that the learner reads complete records.
    Same record layout as LearnerGather; `gathered()` is valid on the learner rank.

    Write-after-read protection (a peer's step k+1 kernel must not store into records the learner is still reading
    from step k):
      * double_buffer=True (default while both buffers stay under 1 GiB): two receive buffers used by step parity. The
        views returned for step k stay valid until the call for step k+1 returns on this stream: a peer can only
        start step k+2 (same buffer) after barrier k+1, which the learner's stream reaches after its stream-ordered
        reads of step k.
      * double_buffer=False (large records, e.g. the synthetic 1000-locale scenario at 3.9 GB per rank): one buffer
        and a second barrier, `release()`, which every rank calls after the learner's reads are enqueued and before
        the next step is launched. Its cost is part of every benchmark that uses this mode."""

    def __init__(self, eng, world: int, learner: int = 0, group=None, double_buffer: bool | None = None):
        import torch.distributed._symmetric_memory as symm
        self.layout = PackedLayout(eng.n_envs, eng.obs_dim, eng.dtype)
        self.world, self.learner, self.rank = world, learner, dist.get_rank(group)
        n = self.layout.nbytes
        if double_buffer is None:
            double_buffer = 2 * world * n <= (1 << 30)
        self.nbuf = 2 if double_buffer else 1
        self.recv = symm.empty(self.nbuf * world * n, dtype=torch.uint8, device=eng.device)
        self.recv.zero_()
        self.hdl = symm.rendezvous(self.recv, group if group is not None else dist.group.WORLD)
        peer = self.hdl.get_buffer(learner, (self.nbuf * world * n,), torch.uint8)  # the learner's buffers, mapped here
        self.mine = [self.layout.views(peer[(b * world + self.rank) * n:(b * world + self.rank + 1) * n])
                     for b in range(self.nbuf)]
        self.cur = 0      # buffer the NEXT step writes
        self.prev = 0     # buffer of the last completed gather
        eng.bind_mirror(*self.mine[0])
        self.eng = eng

    def __call__(self, obs=None, rew=None, done=None):
        # the kernel has already stored this rank's records into the learner's buffer: order the reader behind every
        # rank's step kernel (stream-ordered barrier on the signal pads)
        self.hdl.barrier(channel=0)
        self.prev = self.cur
        if self.nbuf == 2:
            self.cur ^= 1
            self.eng.bind_mirror(*self.mine[self.cur])   # pointers only; the next launch picks them up
        return self.gathered() if self.rank == self.learner else None

    def release(self):
        """single-buffer mode: the learner's reads of the last gather are enqueued; peers may overwrite after this"""
        if self.nbuf == 1:
            self.hdl.barrier(channel=1)

    def gathered(self):
        """[(obs, reward, done) of rank r] of the last completed gather, views of the learner's receive buffer (see the
        class docstring for how long they stay valid)"""
        n, b = self.layout.nbytes, self.prev
        return [self.layout.views(self.recv[(b * self.world + r) * n:(b * self.world + r + 1) * n]) for r in range(self.world)]

    last = gathered

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
