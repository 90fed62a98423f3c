"""Batched PPO over the GPU-resident environments (BASELINE.json configs[4]).

The reference trains with SB3 PPO on ONE env through a DummyVecEnv (runner.py:62-155, hyper-parameters from
configs/ppo_default.yaml); SB3 is not in this image and a host round trip per env step would undo the batched
simulator. This is the in-repo learner the north_star asks for: policy/value MLP 256-256 ReLU as
runner.py:116-119, Gaussian policy clipped to the Box(0,1) action space, GAE(gamma, lambda), clipped surrogate,
value and entropy terms, gradient clipping (algorithm as in the reference's in-house optimizer/ppo.py:88-353;
reward scaling as optimizer/env.py:19-36). Everything stays on the device.

Multi-GPU (dist.py): every rank steps its own env shard with a policy replica and keeps its rollout on its own
device; ONE all-gather per update delivers the packed rollout records (obs | action | log-prob | value | reward | done)
of all shards, the learner rank (0) runs the PPO update and broadcasts the weights.

Determinism across layouts: the exploration noise of env e at step t is a function of (seed, update, t, e) only - every
rank draws the noise of ALL envs from the same generator state and keeps its shard's slice - so a W-rank run collects
exactly the rollout a 1-rank run over the same W*N envs collects (tests/test_learner.py), and ranks never duplicate
each other's rollouts.

On a CUDA device the whole rollout (n_steps x [policy forward, epi_step, episode bookkeeping, masked reset]) is captured
once into a CUDA graph and replayed per update: the per-step host work of an eager PyTorch loop is what kept the
learner at a fifth of the simulator's speed in round 1.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch import nn


class ActorCritic(nn.Module):
    def __init__(self, obs_dim: int, act_dim: int, hidden=(256, 256), init_log_std: float = -1.0):
        super().__init__()
        def mlp(out):
            layers, d = [], obs_dim
            for h in hidden:
                layers += [nn.Linear(d, h), nn.ReLU()]
                d = h
            return nn.Sequential(*layers, nn.Linear(d, out))
        self.pi, self.v = mlp(act_dim), mlp(1)
        self.log_std = nn.Parameter(torch.full((act_dim,), init_log_std))

    def dist(self, obs):
        # validate_args=False: the argument checks synchronise with the host (`.all()`), which is illegal inside the
        # CUDA-graph capture of the rollout and wasted time everywhere else
        return torch.distributions.Normal(torch.sigmoid(self.pi(obs)), self.log_std.exp(), validate_args=False)

    def _logp(self, mean, a):
        return (-0.5 * ((a - mean) / self.log_std.exp()) ** 2 - self.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    @torch.no_grad()
    def act(self, obs, deterministic: bool = False, eps=None):
        """eps: standard-normal noise [N, A] supplied by the caller (reproducible across rank layouts), else sampled.
        Plain tensor arithmetic (no torch.distributions object): capturable in a CUDA graph."""
        mean = torch.sigmoid(self.pi(obs))
        if deterministic:
            a = mean
        else:
            a = mean + self.log_std.exp() * (eps if eps is not None else torch.randn_like(mean))
        return a, self._logp(mean, a), self.v(obs).squeeze(-1)

    def evaluate(self, obs, act):
        d = self.dist(obs)
        return d.log_prob(act).sum(-1), d.entropy().sum(-1), self.v(obs).squeeze(-1)


def gae(rew, val, done, last_val, gamma: float, lam: float):
    """rew/val/done: [T, N]; returns (advantages, returns) [T, N] (optimizer/ppo.py:60-86 restated for a batch)."""
    T = rew.shape[0]
    adv = torch.zeros_like(rew)
    nxt, run = last_val, torch.zeros_like(last_val)
    for t in reversed(range(T)):
        nonterm = 1.0 - done[t]
        delta = rew[t] + gamma * nxt * nonterm - val[t]
        run = delta + gamma * lam * nonterm * run
        adv[t] = run
        nxt = val[t]
    return adv, adv + val


class BatchedPPO:
    """PPO on a BatchedEpiEnv-like object: reset() -> obs [N, D]; step(a [N, A]) -> (obs, reward [N], done [N], _)."""

    def __init__(self, env, obs_dim: int, act_dim: int, obs_scale, n_steps: int = 16, n_epochs: int = 4,
                 minibatches: int = 4, gamma: float = 0.999, lam: float = 0.99, clip: float = 0.2,
                 lr: float = 3e-4, vf_coef: float = 0.974, ent_coef: float = 5.3e-8, max_grad_norm: float = 2.0,
                 reward_scale: float = 1e-7, device="cuda", seed: int = 10, gather=None, use_graph: bool | None = None,
                 desync: int = 0, update_mode: str = "learner"):
        self.env, self.dev = env, torch.device(device)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        torch.manual_seed(seed)  # identical replicas on every rank (same init, weights broadcast after every update)
        self.ac = ActorCritic(obs_dim, act_dim).to(self.dev)
        self.use_graph = (self.dev.type == "cuda") if use_graph is None else use_graph
        # capturable: the step counters live on the device, so the optimiser steps can sit inside the update graph (set
        # with and without graphs: both paths then run the same kernels and produce the same weights)
        self.opt = torch.optim.Adam(self.ac.parameters(), lr=lr, capturable=self.dev.type == "cuda")
        self.obs_scale = torch.as_tensor(obs_scale, dtype=torch.float32, device=self.dev)
        self.n_steps, self.n_epochs, self.minibatches = n_steps, n_epochs, minibatches
        self.gamma, self.lam, self.clip = gamma, lam, clip
        self.vf_coef, self.ent_coef, self.max_grad_norm, self.reward_scale = vf_coef, ent_coef, max_grad_norm, reward_scale
        self.gather = gather
        self.obs_dim, self.act_dim = obs_dim, act_dim
        # exploration noise: one generator per learner, same state on every rank, advanced by whole-fleet draws
        self._noise = torch.Generator(device=self.dev).manual_seed(seed + 7919)
        # minibatch permutations of the learner rank only
        self._perm = torch.Generator(device=self.dev).manual_seed(seed + 104729)
        self._obs = None
        self.ep_return = None
        self._buf = None
        assert update_mode in ("learner", "sharded")
        self.update_mode = update_mode
        self.desync = int(desync)  # episode length in gym steps: spread the envs' episode phases over it (0 = off)
        self._graph, self._graph_failed, self._warm = None, False, False
        self._ugraph, self._ugraph_failed, self._uwarm, self._ubuf = None, False, False, None
        # finished episodes of this rank's shard, kept on the device (no host sync inside the rollout)
        self._fin_sum = torch.zeros((), dtype=torch.float64, device=self.dev)
        self._fin_cnt = torch.zeros((), dtype=torch.int64, device=self.dev)

    @property
    def finished_episodes(self) -> tuple[int, float]:
        """(count, mean return) of the episodes finished so far on this rank (one host sync)"""
        n = int(self._fin_cnt)
        return n, (float(self._fin_sum) / n if n else math.nan)

    def _norm(self, obs):
        return obs.to(torch.float32) / self.obs_scale

    def _all(self, x):
        """all-gather along the env axis: x [..., N_local, F...] with the env axis at dim `-2 or 0` flattened by the
        caller; here x is [N_local, ...] and the result [world * N_local, ...] (rank-major = global env order)"""
        if self.world == 1:
            return x
        x = x.contiguous()
        out = torch.empty((self.world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        if x.is_cuda:
            dist.all_gather_into_tensor(out.view(-1), x.view(-1))
        else:
            dist.all_gather(list(out.unbind(0)), x)
        return out.reshape((-1,) + tuple(x.shape[1:]))

    # ---- rollout ----------------------------------------------------------------------------------
    def _alloc(self, n):
        T, D, A, dev = self.n_steps, self.obs_dim, self.act_dim, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        self._buf = dict(O=torch.zeros((T, n, D), **f32), A=torch.zeros((T, n, A), **f32), LP=torch.zeros((T, n), **f32),
                         V=torch.zeros((T, n), **f32), R=torch.zeros((T, n), **f32), D=torch.zeros((T, n), **f32),
                         eps=torch.zeros((T, n, A), **f32), last_v=torch.zeros(n, **f32))

    def _rollout_body(self):
        """n_steps gym steps of this rank's envs into the static rollout buffers; no host synchronisation (this is
        the region a CUDA graph captures)"""
        B = self._buf
        for t in range(self.n_steps):
            o = self._norm(self._obs)
            a, lp, v = self.ac.act(o, eps=B["eps"][t])
            obs, rew, done, _ = self.env.step(a.clamp(0.0, 1.0).contiguous())
            self.ep_return += rew
            d = done.to(torch.bool)
            # episode bookkeeping and the masked restart stay on the device: no `if d.any()` host round trip per step
            self._fin_sum += torch.where(d, self.ep_return, torch.zeros_like(self.ep_return)).sum()
            self._fin_cnt += d.sum()
            self.ep_return.copy_(torch.where(d, torch.zeros_like(self.ep_return), self.ep_return))
            B["O"][t].copy_(o); B["A"][t].copy_(a); B["LP"][t].copy_(lp); B["V"][t].copy_(v)
            B["R"][t].copy_(rew.to(torch.float32) * self.reward_scale)  # before the restart: reward aliases an engine buffer
            B["D"][t].copy_(d.to(torch.float32))
            self.env.reset(done)  # restarts the envs whose flag is set (a launch that touches nothing otherwise)
            self._obs.copy_(self.env.epi.obs)
        B["last_v"].copy_(self.ac.act(self._norm(self._obs), deterministic=True)[2])

    def collect(self):
        """n_steps gym steps of every env; returns the rollout as [T, N_total, ...] tensors on every rank."""
        if self._obs is None:
            self._obs = self.env.reset().clone()
            n = self._obs.shape[0]
            self.ep_return = torch.zeros(n, dtype=torch.float64, device=self.dev)
            self._alloc(n)
            if self.desync > 1:
                self._desynchronise(n)
        B, n = self._buf, self._obs.shape[0]
        # the fleet's noise for this rollout, this rank's slice of it (see the module docstring)
        eps = torch.randn((self.n_steps, self.world * n, self.act_dim), generator=self._noise, device=self.dev)
        B["eps"].copy_(eps[:, self.rank * n:(self.rank + 1) * n])
        if self.use_graph and not self._graph_failed and self._warm:
            if self._graph is None:
                self._capture()
            if self._graph is not None:
                self._graph.replay()
            else:
                self._rollout_body()
        else:
            self._rollout_body()  # the first rollout runs eagerly: it is also the warm-up a graph capture needs
            self._warm = True
        if self.world == 1 or self.update_mode == "sharded":
            return tuple(B[k] for k in ("O", "A", "LP", "V", "R", "D", "last_v"))
        # ONE exchange per update: the packed rollout record of every shard (rank-major = global env order)
        T, D, A = self.n_steps, self.obs_dim, self.act_dim
        rec = torch.cat([B["O"], B["A"], B["LP"][..., None], B["V"][..., None], B["R"][..., None], B["D"][..., None]], -1)
        rec = torch.cat([rec.reshape(-1), B["last_v"]])
        allr = self._all(rec[None])  # [world, T*n*(D+A+4) + n]
        body = allr[:, :T * n * (D + A + 4)].reshape(self.world, T, n, D + A + 4).permute(1, 0, 2, 3).reshape(T, self.world * n, D + A + 4)
        last_v = allr[:, T * n * (D + A + 4):].reshape(-1)
        return (body[..., :D], body[..., D:D + A], body[..., D + A], body[..., D + A + 1], body[..., D + A + 2],
                body[..., D + A + 3], last_v)

    def _desynchronise(self, n):
        """All envs of a fleet start on day 0, so a rollout of n_steps gym steps sees ONE phase of the epidemic (calm, peak,
        tail) in every env and consecutive updates see different ones: the advantage baseline chases the phase instead
        of the policy. Before the first rollout, env e (global index) is advanced (desync - 1 - e mod desync) gym steps
        with uniform random actions: episode phases are then spread evenly and every rollout covers a whole episode's
        worth of them. Runs once; the noise comes from the fleet generator (same on every rank layout)."""
        L = self.desync
        idx = torch.arange(self.rank * n, (self.rank + 1) * n, device=self.dev) % L
        for t in range(L - 1):
            u = torch.rand((self.world * n, self.act_dim), generator=self._noise, device=self.dev)[self.rank * n:(self.rank + 1) * n]
            self.env.step(u.contiguous())
            self.env.reset((idx > t).to(torch.uint8))  # envs whose offset is not reached yet start over
        self._obs.copy_(self.env.epi.obs)
        self.ep_return.zero_()

    def _capture(self):
        """capture the rollout into a CUDA graph (the first, eager rollout was the warm-up: allocator, cuBLAS handles,
        specialised kernels are in place); capturing does not execute, the caller replays. Any failure leaves the eager
        path in charge."""
        try:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._rollout_body()
            self._graph = g
        except Exception as e:  # noqa: BLE001
            import sys
            print(f"[epi-b200] rollout graph capture failed ({type(e).__name__}: {e}); eager rollout", file=sys.stderr)
            self._graph, self._graph_failed = None, True
            torch.cuda.synchronize(self.dev)

    def update(self, batch):
        """`learner` mode (the north_star's wording): the gathered rollout of the whole fleet is optimised on rank 0 and
        the weights are broadcast. `sharded` mode: every rank optimises on ITS OWN shard's rollout (no rollout exchange
        at all), the flattened gradients of every minibatch step are averaged with one all-reduce, and the advantage
        normalisation uses fleet-wide moments - the update then costs what it costs on one GPU however many GPUs step
        envs, and the replicas stay bit-identical without a broadcast (same averaged gradients, same optimiser state).

        On a CUDA device the whole update (GAE, advantage normalisation, n_epochs x minibatches optimiser steps with
        their gradient all-reduces) is ONE CUDA graph over static buffers, captured after the first (eager) update and
        replayed afterwards; only the minibatch permutations are drawn outside it."""
        sharded = self.world > 1 and self.update_mode == "sharded"
        stats = {}
        if self.rank == 0 or sharded:
            U = self._update_buffers(batch)
            for e in range(self.n_epochs):
                U["perm"][e].copy_(torch.randperm(U["perm"].shape[1], device=self.dev, generator=self._perm))
            if self.use_graph and not self._ugraph_failed and self._uwarm:
                if self._ugraph is None:
                    self._capture_update(U)
                if self._ugraph is not None:
                    self._ugraph.replay()
                else:
                    self._update_body(U)
            else:
                self._update_body(U)  # the first update runs eagerly: warm-up of cuBLAS, Adam state, NCCL
                self._uwarm = True
            pg, vl, ent = U["stats"].tolist()  # the update's one host synchronisation
            stats = {"pg_loss": pg, "v_loss": vl, "entropy": ent}
        if self.world > 1 and not sharded:  # the learner's weights go back to every shard's replica
            flat = torch.cat([p.data.reshape(-1) for p in self.ac.parameters()])
            dist.broadcast(flat, 0)
            off = 0
            for p in self.ac.parameters():
                p.data.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        return stats

    def _update_buffers(self, batch):
        """static inputs of the update: the rollout tensors themselves when they already are this learner's static
        rollout buffers (one rank, or `sharded`), else (the gathered fleet rollout on the learner rank) copies of them"""
        names = ("O", "A", "LP", "V", "R", "D", "last_v")
        if self._ubuf is None:
            own = self._buf is not None and all(x is self._buf[k] for k, x in zip(names, batch))
            U = {k: (x if own else torch.empty_like(x, memory_format=torch.contiguous_format)) for k, x in zip(names, batch)}
            n = batch[4].numel()
            U["perm"] = torch.zeros((self.n_epochs, n), dtype=torch.int64, device=self.dev)
            U["stats"] = torch.zeros(3, dtype=torch.float32, device=self.dev)
            U["own"] = own
            self._ubuf = U
        U = self._ubuf
        if not U["own"]:
            for k, x in zip(names, batch):
                U[k].copy_(x)
        return U

    def _update_body(self, U):
        """GAE, advantage normalisation and the optimiser steps over the static buffers; no host synchronisation (this is
        the region the update graph captures)"""
        sharded = self.world > 1 and self.update_mode == "sharded"
        adv, ret = gae(U["R"], U["V"], U["D"], U["last_v"], self.gamma, self.lam)
        O, A, LP, adv, ret = (x.reshape(-1, *x.shape[2:]) for x in (U["O"], U["A"], U["LP"], adv, ret))
        if sharded:
            m = torch.stack([adv.sum(dtype=torch.float64), (adv.double() ** 2).sum(),
                             torch.full((), float(adv.numel()), dtype=torch.float64, device=adv.device)])
            dist.all_reduce(m)
            mean = m[0] / m[2]
            std = ((m[1] - m[2] * mean ** 2) / (m[2] - 1)).clamp_min(0).sqrt()
            adv = ((adv - mean.float()) / (std.float() + 1e-8))
        else:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        params = list(self.ac.parameters())
        pi_params = list(self.ac.pi.parameters()) + [self.ac.log_std]
        v_params = list(self.ac.v.parameters())
        for e in range(self.n_epochs):
            for idx in U["perm"][e].chunk(self.minibatches):
                lp, ent, v = self.ac.evaluate(O[idx], A[idx])
                ratio = (lp - LP[idx]).exp()
                pg = -torch.min(ratio * adv[idx], ratio.clamp(1 - self.clip, 1 + self.clip) * adv[idx]).mean()
                vl = ((v - ret[idx]) ** 2).mean()
                ent = ent.mean()
                loss = pg + self.vf_coef * vl - self.ent_coef * ent
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                if sharded:
                    flat = torch.cat([p.grad.reshape(-1) for p in params])
                    dist.all_reduce(flat)
                    flat /= self.world
                    off = 0
                    for p in params:
                        p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                        off += p.numel()
                # policy and value networks are separate: each is clipped on its own norm (one joint norm lets the
                # value loss, large while the critic is still fitting returns of O(10^2), throttle the policy step)
                nn.utils.clip_grad_norm_(pi_params, self.max_grad_norm)
                nn.utils.clip_grad_norm_(v_params, self.max_grad_norm)
                self.opt.step()
        U["stats"].copy_(torch.stack([pg.detach(), vl.detach(), ent.detach()]))

    def _capture_update(self, U):
        """capture the update into a CUDA graph (the first, eager update was the warm-up). Gradients are dropped first so
        that the captured backward allocates them from the graph's pool. Any failure leaves the eager path in charge."""
        try:
            torch.cuda.synchronize(self.dev)
            self.opt.zero_grad(set_to_none=True)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._update_body(U)
            self._ugraph = g
        except Exception as e:  # noqa: BLE001
            import sys
            print(f"[epi-b200] update graph capture failed ({type(e).__name__}: {e}); eager update", file=sys.stderr)
            self._ugraph, self._ugraph_failed = None, True
            torch.cuda.synchronize(self.dev)
            self.opt.zero_grad(set_to_none=True)

    def train(self, n_updates: int, log=None):
        hist = []
        for u in range(n_updates):
            batch = self.collect()
            st = self.update(batch)
            mr = batch[4].sum(0).mean()
            if self.world > 1 and self.update_mode == "sharded":
                mr = mr.clone()
                dist.all_reduce(mr)
                mr = mr / self.world
            mean_r = float(mr) / self.reward_scale / self.n_steps
            n_ep, mean_ep = self.finished_episodes
            st.update(update=u, mean_step_reward=mean_r, episodes=n_ep, mean_episode_return=mean_ep)
            hist.append(st)
            if log and self.rank == 0:
                log(st)
        return hist


# ----------------------------------------------------------------------------------------------
# SAC (runner.py --algo sac; algorithm as the reference's in-house optimizer/sac.py:44-341: replay buffer,
# twin Q functions with Polyak-averaged targets, squashed Gaussian policy; hyper-parameters as
# configs/sac_default.yaml) restated for a batch of GPU-resident envs: the replay buffer is a ring of device
# tensors fed N transitions per gym step, actions are squashed into the env's Box(0, 1) action space.
# ----------------------------------------------------------------------------------------------
def _mlp(inp: int, out: int, hidden=(256, 256)):
    layers, d = [], inp
    for h in hidden:
        layers += [nn.Linear(d, h), nn.ReLU()]
        d = h
    return nn.Sequential(*layers, nn.Linear(d, out))


class SquashedGaussianActor(nn.Module):
    LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0

    def __init__(self, obs_dim: int, act_dim: int, hidden=(256, 256)):
        super().__init__()
        self.net = _mlp(obs_dim, 2 * act_dim, hidden)

    def forward(self, obs, deterministic: bool = False, with_logprob: bool = True):
        mu, log_std = self.net(obs).chunk(2, -1)
        log_std = log_std.clamp(self.LOG_STD_MIN, self.LOG_STD_MAX)
        u = mu if deterministic else mu + log_std.exp() * torch.randn_like(mu)
        a = 0.5 * (torch.tanh(u) + 1.0)  # Box(0, 1)
        logp = None
        if with_logprob:
            # N(u; mu, std) with the change of variables a = (tanh(u) + 1) / 2:  da/du = (1 - tanh(u)^2) / 2
            logp = (-0.5 * ((u - mu) / log_std.exp()) ** 2 - log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
            logp = logp - (2.0 * (math.log(2.0) - u - nn.functional.softplus(-2.0 * u)) - math.log(2.0)).sum(-1)
        return a, logp


class ReplayRing:
    """Fixed-size ring of transitions in device memory; `store` takes a whole batch of envs at once."""

    def __init__(self, obs_dim: int, act_dim: int, size: int, device):
        self.o = torch.zeros((size, obs_dim), device=device)
        self.o2 = torch.zeros((size, obs_dim), device=device)
        self.a = torch.zeros((size, act_dim), device=device)
        self.r = torch.zeros(size, device=device)
        self.d = torch.zeros(size, device=device)
        self.size, self.ptr, self.n = size, 0, 0

    def store(self, o, a, r, o2, d):
        m = o.shape[0]
        if m > self.size:
            o, a, r, o2, d = (x[-self.size:] for x in (o, a, r, o2, d))
            m = self.size
        idx = (self.ptr + torch.arange(m, device=self.o.device)) % self.size
        self.o[idx], self.a[idx], self.r[idx], self.o2[idx], self.d[idx] = o, a, r, o2, d
        self.ptr = (self.ptr + m) % self.size
        self.n = min(self.size, self.n + m)

    def sample(self, batch: int):
        idx = torch.randint(0, self.n, (batch,), device=self.o.device)
        return self.o[idx], self.a[idx], self.r[idx], self.o2[idx], self.d[idx]


class BatchedSAC:
    """SAC on a BatchedEpiEnv-like object (same env protocol as BatchedPPO). Multi-GPU: every rank steps its shard
    with an actor replica; one all-gather per gym step delivers the packed transitions; rank 0 owns the replay ring,
    runs the gradient steps and broadcasts the actor."""

    def __init__(self, env, obs_dim: int, act_dim: int, obs_scale, gamma: float = 0.9, tau: float = 0.1,
                 lr: float = 3e-3, batch_size: int = 256, buffer_size: int = 1 << 20, learning_starts: int = 0,
                 gradient_steps: int = 4, target_entropy: float | None = None, reward_scale: float = 1e-7,
                 device="cuda", seed: int = 10):
        self.env, self.dev = env, torch.device(device)
        self.obs_scale = torch.as_tensor(obs_scale, dtype=torch.float32, device=self.dev)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        torch.manual_seed(seed)  # identical networks on every rank ...
        self.actor = SquashedGaussianActor(obs_dim, act_dim).to(self.dev)
        self.q1, self.q2 = _mlp(obs_dim + act_dim, 1).to(self.dev), _mlp(obs_dim + act_dim, 1).to(self.dev)
        self.q1_t, self.q2_t = _mlp(obs_dim + act_dim, 1).to(self.dev), _mlp(obs_dim + act_dim, 1).to(self.dev)
        self.q1_t.load_state_dict(self.q1.state_dict()); self.q2_t.load_state_dict(self.q2.state_dict())
        for p in list(self.q1_t.parameters()) + list(self.q2_t.parameters()):
            p.requires_grad_(False)
        self.log_alpha = torch.zeros((), device=self.dev, requires_grad=True)
        self.target_entropy = float(-act_dim if target_entropy is None else target_entropy)
        self.opt_pi = torch.optim.Adam(self.actor.parameters(), lr=lr)
        self.opt_q = torch.optim.Adam(list(self.q1.parameters()) + list(self.q2.parameters()), lr=lr)
        self.opt_alpha = torch.optim.Adam([self.log_alpha], lr=lr)
        self.gamma, self.tau, self.batch_size, self.gradient_steps = gamma, tau, batch_size, gradient_steps
        self.learning_starts, self.reward_scale = learning_starts, reward_scale
        self.ring = ReplayRing(obs_dim, act_dim, buffer_size, self.dev) if self.rank == 0 else None
        self._obs, self.total_steps = None, 0
        self.ep_return = None
        self._fin_sum = torch.zeros((), dtype=torch.float64, device=self.dev)
        self._fin_cnt = torch.zeros((), dtype=torch.int64, device=self.dev)

    def _norm(self, obs):
        return obs.to(torch.float32) / self.obs_scale

    def _all(self, x):
        if self.world == 1:
            return x
        out = [torch.empty_like(x) for _ in range(self.world)]
        dist.all_gather(out, x.contiguous())
        return torch.cat(out, 0)

    @torch.no_grad()
    def collect_step(self):
        """one gym step of every env; the transitions of all shards land in the learner's ring"""
        if self._obs is None:
            self._obs = self.env.reset().clone()
            self.ep_return = torch.zeros(self._obs.shape[0], dtype=torch.float64, device=self.dev)
        o = self._norm(self._obs)
        if self.total_steps < self.learning_starts:
            a = torch.rand((o.shape[0], self.actor.net[-1].out_features // 2), device=self.dev)
        else:
            a, _ = self.actor(o, with_logprob=False)
        obs, rew, done, _ = self.env.step(a.contiguous())
        o2 = self._norm(obs)  # the terminal observation, before any reset
        self.ep_return += rew
        d = done.to(torch.bool)
        self._fin_sum += torch.where(d, self.ep_return, torch.zeros_like(self.ep_return)).sum()
        self._fin_cnt += d.sum()
        self.ep_return = torch.where(d, torch.zeros_like(self.ep_return), self.ep_return)
        r = rew.to(torch.float32) * self.reward_scale
        mean_r = rew.mean()
        self.env.reset(done)  # masked restart, no host round trip
        obs = self.env.epi.obs
        rec = self._all(torch.cat([o, a, r[:, None], o2, d.to(torch.float32)[:, None]], 1))
        if self.rank == 0:
            nd, na = o.shape[1], a.shape[1]
            self.ring.store(rec[:, :nd], rec[:, nd:nd + na], rec[:, nd + na], rec[:, nd + na + 1:2 * nd + na + 1],
                            rec[:, 2 * nd + na + 1])
        self._obs = obs.clone()
        self.total_steps += rec.shape[0]
        return mean_r

    def update(self):
        stats = {}
        if self.rank == 0 and self.ring.n >= self.batch_size and self.total_steps >= self.learning_starts:
            for _ in range(self.gradient_steps):
                o, a, r, o2, d = self.ring.sample(self.batch_size)
                alpha = self.log_alpha.exp().detach()
                with torch.no_grad():
                    a2, logp2 = self.actor(o2)
                    oa2 = torch.cat([o2, a2], -1)
                    q_t = torch.min(self.q1_t(oa2), self.q2_t(oa2)).squeeze(-1)
                    backup = r + self.gamma * (1.0 - d) * (q_t - alpha * logp2)
                oa = torch.cat([o, a], -1)
                loss_q = ((self.q1(oa).squeeze(-1) - backup) ** 2).mean() + ((self.q2(oa).squeeze(-1) - backup) ** 2).mean()
                self.opt_q.zero_grad(set_to_none=True)
                loss_q.backward()
                self.opt_q.step()
                pi, logp = self.actor(o)
                op = torch.cat([o, pi], -1)
                loss_pi = (alpha * logp - torch.min(self.q1(op), self.q2(op)).squeeze(-1)).mean()
                self.opt_pi.zero_grad(set_to_none=True)
                loss_pi.backward()
                self.opt_pi.step()
                loss_alpha = -(self.log_alpha * (logp.detach() + self.target_entropy).mean())
                self.opt_alpha.zero_grad(set_to_none=True)
                loss_alpha.backward()
                self.opt_alpha.step()
                with torch.no_grad():
                    for net, tgt in ((self.q1, self.q1_t), (self.q2, self.q2_t)):
                        for p, pt in zip(net.parameters(), tgt.parameters()):
                            pt.mul_(1.0 - self.tau).add_(p, alpha=self.tau)
            stats = {"q_loss": float(loss_q.detach()), "pi_loss": float(loss_pi.detach()), "alpha": float(self.log_alpha.detach().exp())}
        if self.world > 1:
            flat = torch.cat([p.data.reshape(-1) for p in self.actor.parameters()])
            dist.broadcast(flat, 0)
            off = 0
            for p in self.actor.parameters():
                p.data.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        return stats

    def train(self, n_steps: int, log=None):
        hist = []
        for s in range(n_steps):
            mean_r = self.collect_step()
            st = self.update()
            st.update(step=s, mean_step_reward=float(mean_r), episodes=int(self._fin_cnt))
            hist.append(st)
            if log and self.rank == 0:
                log(st)
        return hist
