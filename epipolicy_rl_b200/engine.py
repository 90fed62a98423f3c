"""BatchedEpidemic — host-side mirror of the reference's `Epidemic` (core/epidemic.py:34-116) for N
independent instances of one scenario, driving libepi_b200.so through its C-ABI. PyTorch is used only
for device buffers and streams."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import lib as _lib
from .desc import EpiDesc

PERIOD = 7  # epipolicy_environment.py:1


class BatchedEpidemic:
    """N copies of one scenario on one GPU.

    precision: "fp64" = parity mode (mirrors the reference's float32/float64 roundings, 1e-9),
               "fp32" = throughput mode (state and stages in float32, 1e-4).
    """

    def __init__(self, desc: EpiDesc, n_envs: int, device: int | str | torch.device = 0, precision: str = "fp64"):
        self.L = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedEpidemic needs a CUDA device (no CPU fallback)")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.device = dev
        self.desc = desc
        self.n_envs = int(n_envs)
        self.precision = {"fp64": _lib.EPI_FP64, "fp32": _lib.EPI_FP32}[precision]
        self.dtype = torch.float64 if precision == "fp64" else torch.float32
        self.np_dtype = np.float64 if precision == "fp64" else np.float32
        self._c = desc.as_c()
        h = ctypes.c_void_p()
        rc = self.L.epi_create(ctypes.byref(self._c), self.n_envs, dev.index or 0, self.precision, ctypes.byref(h))
        if rc != 0:
            raise _lib.EpiError(rc, self.L.epi_last_error(None).decode())
        self._h = h
        self.info = _lib.EpiInfo()
        self._ck(self.L.epi_info(self._h, ctypes.byref(self.info)))
        self.obs_dim, self.action_dim = self.info.obs_dim, self.info.action_dim
        if self.info.team_kind in (2, 3):
            # scenario-specialised build of the cell kernel (spec.py); the interpreted kernel runs if absent
            from . import spec as _spec
            path = _spec.ensure(desc, self.precision)
            if path is not None:
                rc = self.L.epi_spec_attach(self._h, path.encode())
                if rc != -2:  # EPI_E_UNSUPPORTED: a kernel kind forced with EPI_TEAM has no specialised build
                    self._ck(rc)
                self._ck(self.L.epi_info(self._h, ctypes.byref(self.info)))
        with torch.cuda.device(dev):
            self.obs = torch.empty((self.n_envs, self.obs_dim), dtype=self.dtype, device=dev)
            self.reward = torch.zeros(self.n_envs, dtype=torch.float64, device=dev)
            self.done = torch.zeros(self.n_envs, dtype=torch.uint8, device=dev)
        self.reset()

    # ---------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise _lib.EpiError(rc, self.L.epi_last_error(self._h).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None):
            self.L.epi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def bind_outputs(self, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor) -> None:
        """Let the kernel epilogue write obs / reward / done into caller-owned device buffers (e.g. the
        send buffer of the learner all-gather, dist.py)."""
        assert obs.shape == (self.n_envs, self.obs_dim) and obs.dtype == self.dtype and obs.is_contiguous()
        assert reward.shape == (self.n_envs,) and reward.dtype == torch.float64
        assert done.shape == (self.n_envs,) and done.dtype == torch.uint8
        obs.copy_(self.obs); reward.copy_(self.reward); done.copy_(self.done)
        self.obs, self.reward, self.done = obs, reward, done

    def bind_mirror(self, obs: torch.Tensor | None, reward: torch.Tensor | None, done: torch.Tensor | None) -> None:
        """Second destination of every following step's results, written by the kernel epilogue itself: tensors that
        alias the learner rank's receive buffer (peer memory, dist.PeerGather). None unbinds."""
        if obs is not None:
            assert obs.shape == (self.n_envs, self.obs_dim) and obs.dtype == self.dtype and obs.is_contiguous()
            assert reward.shape == (self.n_envs,) and reward.dtype == torch.float64
            assert done.shape == (self.n_envs,) and done.dtype == torch.uint8
        self._mirror = (obs, reward, done)  # keep the mappings alive
        self._ck(self.L.epi_bind_mirror(self._h, *(None if t is None else t.data_ptr() for t in (obs, reward, done))))

    # ---------------------------------------------------------------------------------
    def reset(self, mask: torch.Tensor | None = None) -> torch.Tensor:
        """Epidemic.reset (epidemic.py:104-106) for all envs, or those where mask != 0."""
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        self._ck(self.L.epi_reset(self._h, m.data_ptr() if m is not None else None, self.obs.data_ptr(), self._stream()))
        return self.obs

    def step(self, actions: torch.Tensor, n_days: int = PERIOD):
        """EpiEnv.step for the batch: actions [N, A] float32 in [0, 1] on the device.
        Returns (obs [N, C*L*G], reward [N] float64, done [N] uint8) — views of engine-owned tensors."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if a.shape != (self.n_envs, self.action_dim):
            raise ValueError(f"actions must have shape {(self.n_envs, self.action_dim)}, got {tuple(a.shape)}")
        self._ck(self.L.epi_step(self._h, a.data_ptr(), int(n_days), self.obs.data_ptr(), self.reward.data_ptr(),
                                 self.done.data_ptr(), self._stream()))
        return self.obs, self.reward, self.done

    def step_plan(self, cpv: torch.Tensor, active: torch.Tensor | None, schedule_programs: bool = True, n_days: int = 1):
        """Epidemic.step with explicit Acts (full control-parameter vectors + active flags)."""
        c = cpv.to(device=self.device, dtype=torch.float32).contiguous()
        a = None if active is None else active.to(device=self.device, dtype=torch.uint8).contiguous()
        self._ck(self.L.epi_step_plan(self._h, c.data_ptr(), a.data_ptr() if a is not None else None,
                                      1 if schedule_programs else 0, int(n_days), self.obs.data_ptr(),
                                      self.reward.data_ptr(), self.done.data_ptr(), self._stream()))
        return self.obs, self.reward, self.done

    # ---------------------------------------------------------------------------------
    def plan_tables(self) -> dict:
        """static tables of the compiled scenario (epi_get_plan): positions of the live accumulators in the reference's
        flat observable, locale classes, per-class default parameters / facility_timespent and their multiplier classes"""
        d, n_lc = self.desc, self.info.n_lc
        spec = {"acc_flat": (0, (self.info.n_acc,), np.int32), "lclass": (1, (d.L,), np.int32),
                "mp0": (2, (n_lc, d.P, d.F, d.G), np.float32), "mp_cls": (3, (n_lc, d.P, d.F, d.G), np.int32),
                "ft0": (4, (n_lc, d.F, d.G), np.float32), "ft_cls": (5, (n_lc, d.F, d.G), np.int32)}
        out = {}
        for k, (code, shape, dt) in spec.items():
            a = np.empty(shape, dt)
            self._ck(self.L.epi_get_plan(self._h, code, a.ctypes.data, a.nbytes))
            out[k] = a
        return out

    def state_dev(self, what: str) -> torch.Tensor:
        """stream-ordered device copy of a piece of the state (no host round trip): "X" [N, C*L*G] f64, "cost" [N, I*L] f64,
        "mult" [N, n_cls] f32, "acc" [N, n_acc, L*G] in the engine's precision (the live cumulative components)"""
        d, N = self.desc, self.n_envs
        code, shape, dt = {"X": (0, (N, self.obs_dim), torch.float64), "cost": (1, (N, d.I * d.L), torch.float64),
                           "mult": (7, (N, self.info.n_cls), torch.float32),
                           "acc": (10, (N, self.info.n_acc, d.L * d.G), self.dtype)}[what]
        t = torch.empty(shape, dtype=dt, device=self.device)
        self._ck(self.L.epi_get_state_dev(self._h, code, t.data_ptr(), t.numel() * t.element_size(), self._stream()))
        return t

    def run(self, cpv: torch.Tensor, active: torch.Tensor | None = None, record_obs: bool = True,
            schedule_programs: bool = True, report: bool = False, one_launch: bool = True):
        """Epidemic.run (core/epidemic.py:118-134) for a batch of what-if schedules: reset, then the whole schedule in
        ONE launch (epi_run_plan) with day t's Acts taken from row t. cpv: [T, N, NCP] (or [T, NCP], broadcast to all envs) full
        control-parameter vectors; active: [T, N, I] / [T, I] flags of the interventions that have an Act on
        day t (None: all). Returns {"reward": [T, N] float64, "done": [N] uint8, "obs": [T, N, C*L*G]} on the device
        (obs only if record_obs). report=True adds "report": the reference Reporter's quantities for every env -
        incidences, average infectious duration, windowed statistics, instantaneous / basic / case reproduction numbers
        per reporting location (core/reporter.py:45-204) - computed on the device from stream-ordered copies of the
        state, the live cumulative components and the class multipliers of every day (reporter_device.py)."""
        c = cpv.to(device=self.device, dtype=torch.float32)
        T = c.shape[0]
        if c.dim() == 2:
            c = c[:, None, :].expand(T, self.n_envs, c.shape[-1])
        a = None
        if active is not None:
            a = active.to(device=self.device, dtype=torch.uint8)
            if a.dim() == 2:
                a = a[:, None, :].expand(T, self.n_envs, a.shape[-1])
        if c.shape[1] != self.n_envs or (a is not None and a.shape[:2] != c.shape[:2]):
            raise ValueError("schedule must be [T, N, NCP] / [T, NCP] with matching active flags")
        self.reset()
        if not one_launch:
            return self._run_daily(c, a, T, record_obs, schedule_programs, report)
        # ONE launch for the whole schedule (epi_run_plan): day d of the launch reads row d of the schedule and emits the
        # day's end state and reward - and, for the reporter, the state in double, the live cumulative components and
        # the class multipliers; no per-day launch, no per-day copies
        from .lib import EpiRunOut
        c = c.contiguous()
        a = None if a is None else a.contiguous()
        N, d = self.n_envs, self.desc
        rew = torch.zeros((T, N), dtype=torch.float64, device=self.device)
        obs = torch.zeros((T, N, self.obs_dim), dtype=self.dtype, device=self.device) if record_obs else None
        rep = x_days = acc_days = mult_days = None
        if report:
            from .reporter_device import DeviceReporter
            rep = DeviceReporter(self.desc, self.plan_tables(), self.device)
            rep.record(self.state_dev("X"), self.state_dev("acc"), self.state_dev("mult"))  # the state after reset
            x_days = torch.zeros((T, N, self.obs_dim), dtype=torch.float64, device=self.device)
            acc_days = torch.zeros((T, N, self.info.n_acc, d.L * d.G), dtype=self.dtype, device=self.device)
            mult_days = torch.zeros((T, N, self.info.n_cls), dtype=torch.float32, device=self.device)
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        days = EpiRunOut(ptr(obs), ptr(rew), ptr(x_days), ptr(acc_days), ptr(mult_days))
        import ctypes
        self._ck(self.L.epi_run_plan(self._h, c.data_ptr(), ptr(a), 1 if schedule_programs else 0, int(T), ctypes.byref(days),
                                     self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(), self._stream()))
        out = {"reward": rew, "done": self.done}
        if rep is not None:
            for t in range(T):
                rep.record(x_days[t], acc_days[t], mult_days[t])
            out["report"] = rep.compute()
        if record_obs:
            out["obs"] = obs
        return out

    def _run_daily(self, c, a, T, record_obs, schedule_programs, report):
        """run(one_launch=False): one epi_step_plan launch per simulated day, results and the reporter's inputs copied
        between the launches (stream-ordered device copies) - the round-1 path, kept as the A/B of epi_run_plan"""
        rep = None
        if report:
            from .reporter_device import DeviceReporter
            rep = DeviceReporter(self.desc, self.plan_tables(), self.device)
            rep.record(self.state_dev("X"), self.state_dev("acc"), self.state_dev("mult"))
        rew = torch.empty((T, self.n_envs), dtype=torch.float64, device=self.device)
        obs = torch.empty((T, self.n_envs, self.obs_dim), dtype=self.dtype, device=self.device) if record_obs else None
        for t in range(T):
            o, r, _ = self.step_plan(c[t], None if a is None else a[t], schedule_programs, 1)
            rew[t].copy_(r)
            if record_obs:
                obs[t].copy_(o)
            if rep is not None:
                rep.record(self.state_dev("X"), self.state_dev("acc"), self.state_dev("mult"))
        out = {"reward": rew, "done": self.done}
        if rep is not None:
            out["report"] = rep.compute()
        if record_obs:
            out["obs"] = obs
        return out

    def step_host(self, actions: np.ndarray, n_days: int = PERIOD):
        """The same step through host (numpy) buffers: H2D, launch, D2H inside the call."""
        a = np.ascontiguousarray(actions, np.float32)
        if a.shape != (self.n_envs, self.action_dim):
            raise ValueError(f"actions must have shape {(self.n_envs, self.action_dim)}")
        if not hasattr(self, "_h_obs"):
            # numpy views of the library's pinned staging buffers: results are read in place (no extra copies);
            # they are overwritten by the next step_host call
            pa, po, pr, pd = (ctypes.c_void_p() for _ in range(4))
            self._ck(self.L.epi_host_buffers(self._h, ctypes.byref(pa), ctypes.byref(po), ctypes.byref(pr), ctypes.byref(pd)))

            def view(ptr, shape, dtype):
                n = int(np.prod(shape)) * np.dtype(dtype).itemsize
                return np.frombuffer((ctypes.c_char * n).from_address(ptr.value), dtype=dtype).reshape(shape)
            self._h_act = view(pa, (self.n_envs, max(self.action_dim, 1)), np.float32)
            self._h_obs = view(po, (self.n_envs, self.obs_dim), self.np_dtype)
            self._h_rew = view(pr, (self.n_envs,), np.float64)
            self._h_done = view(pd, (self.n_envs,), np.uint8)
        if self.action_dim:
            np.copyto(self._h_act[:, :self.action_dim], a)
        self._ck(self.L.epi_step_host(self._h, self._h_act.ctypes.data, int(n_days), self._h_obs.ctypes.data,
                                      self._h_rew.ctypes.data, self._h_done.ctypes.data))
        return self._h_obs, self._h_rew, self._h_done

    # ---------------------------------------------------------------------------------
    def get_state(self, what: str) -> np.ndarray:
        d, N = self.desc, self.n_envs
        spec = {
            "X": (0, (N, d.C, d.L, d.G), np.float64), "cost": (1, (N, d.I, d.L), np.float64),
            "day": (2, (N,), np.int32), "flat": (3, (N, d.n_flat), np.float64), "trace": (4, (N, 4), np.int32),
            "last_err": (5, (N,), np.float64), "cost_rate": (6, (N, d.I, d.L), np.float64),
            "mult": (7, (N, self.info.n_cls), np.float32), "moves": (8, (N, self.info.n_slot, d.L, d.G) if self.info.mv_len == self.info.n_slot * d.L * d.G
                                                                          else (N, self.info.mv_len), np.float32),
            "attempts": (9, (N, 64, 2), np.float64),
        }[what]
        out = np.empty(spec[1], spec[2])
        self._ck(self.L.epi_get_state(self._h, spec[0], out.ctypes.data, out.nbytes))
        return out

    def set_state(self, what: str, value: np.ndarray) -> None:
        code, dt = {"X": (0, np.float64), "cost": (1, np.float64), "day": (2, np.int32), "cost_rate": (6, np.float64),
                    "mult": (7, np.float32), "moves": (8, np.float32)}[what]
        v = np.ascontiguousarray(value, dt)
        self._ck(self.L.epi_set_state(self._h, code, v.ctypes.data, v.nbytes))

    @property
    def launches(self) -> int:
        return int(self.L.epi_launch_count(self._h))
