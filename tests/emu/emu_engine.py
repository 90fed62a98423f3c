"""numpy-buffer driver of the HOST EMULATION build (tests only) — same C-ABI calls as
epipolicy_rl_b200.engine.BatchedEpidemic, with host arrays standing in for device tensors."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from build_emu import build  # noqa: E402

from epipolicy_rl_b200 import lib as _lib  # noqa: E402

_L = None


def emu_lib():
    global _L
    if _L is None:
        _L = _lib.bind(ctypes.CDLL(build()))
    return _L


class EmuEngine:
    def __init__(self, desc, n_envs, precision="fp64"):
        self.L, self.desc, self.n_envs = emu_lib(), desc, n_envs
        self.np_dtype = np.float64 if precision == "fp64" else np.float32
        self._c = desc.as_c()
        h = ctypes.c_void_p()
        rc = self.L.epi_create(ctypes.byref(self._c), n_envs, 0, 0 if precision == "fp64" else 1, ctypes.byref(h))
        if rc != 0:
            raise _lib.EpiError(rc, self.L.epi_last_error(None).decode())
        self._h = h
        self.info = _lib.EpiInfo()
        self.L.epi_info(self._h, ctypes.byref(self.info))
        self.obs = np.zeros((n_envs, self.info.obs_dim), self.np_dtype)
        self.reward = np.zeros(n_envs)
        self.done = np.zeros(n_envs, np.uint8)

    def _ck(self, rc):
        if rc != 0:
            raise _lib.EpiError(rc, self.L.epi_last_error(self._h).decode())

    def close(self):
        if self._h:
            self.L.epi_destroy(self._h)
            self._h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._ck(self.L.epi_reset(self._h, m.ctypes.data if m is not None else None, self.obs.ctypes.data, None))
        return self.obs

    def step(self, actions, n_days=7):
        a = np.ascontiguousarray(actions, np.float32)
        self._ck(self.L.epi_step(self._h, a.ctypes.data, n_days, self.obs.ctypes.data, self.reward.ctypes.data,
                                 self.done.ctypes.data, None))
        return self.obs, self.reward, self.done

    def step_plan(self, cpv, active, schedule_programs=True, n_days=1):
        c = np.ascontiguousarray(cpv, np.float32)
        a = None if active is None else np.ascontiguousarray(active, np.uint8)
        self._ck(self.L.epi_step_plan(self._h, c.ctypes.data, a.ctypes.data if a is not None else None,
                                      1 if schedule_programs else 0, n_days, self.obs.ctypes.data,
                                      self.reward.ctypes.data, self.done.ctypes.data, None))
        return self.obs, self.reward, self.done

    def run_plan(self, cpv_days, active_days, schedule_programs=True):
        """epi_run_plan: [T, N, NCP] / [T, N, I] schedule rows, one launch; returns (obs_days [T, N, C*L*G], reward_days [T, N])"""
        c = np.ascontiguousarray(cpv_days, np.float32)
        a = None if active_days is None else np.ascontiguousarray(active_days, np.uint8)
        T = c.shape[0]
        obs_days = np.zeros((T,) + self.obs.shape, self.obs.dtype)
        rew_days = np.zeros((T, self.n_envs), np.float64)
        from epipolicy_rl_b200.lib import EpiRunOut
        d = self.desc
        self.x_days = np.zeros((T, self.n_envs, d.C * d.L * d.G), np.float64)
        self.acc_days = np.zeros((T, self.n_envs, self.info.n_acc, d.L * d.G), self.obs.dtype)
        self.mult_days = np.zeros((T, self.n_envs, self.info.n_cls), np.float32)
        out = EpiRunOut(obs_days.ctypes.data, rew_days.ctypes.data, self.x_days.ctypes.data, self.acc_days.ctypes.data,
                        self.mult_days.ctypes.data)
        self._ck(self.L.epi_run_plan(self._h, c.ctypes.data, a.ctypes.data if a is not None else None,
                                     1 if schedule_programs else 0, T, ctypes.byref(out),
                                     self.obs.ctypes.data, self.reward.ctypes.data, self.done.ctypes.data, None))
        return obs_days, rew_days

    def acc(self):
        """the live cumulative components [N, n_acc, L*G] (epi_get_state_dev what = 10; "device" memory is host memory here)"""
        d = self.desc
        out = np.empty((self.n_envs, self.info.n_acc, d.L * d.G), self.obs.dtype)
        self._ck(self.L.epi_get_state_dev(self._h, 10, out.ctypes.data, out.nbytes, None))
        return out

    def get_state(self, what):
        d, N = self.desc, self.n_envs
        spec = {"X": (0, (N, d.C, d.L, d.G), np.float64), "cost": (1, (N, d.I, d.L), np.float64),
                "day": (2, (N,), np.int32), "flat": (3, (N, d.n_flat), np.float64), "trace": (4, (N, 4), np.int32),
                "last_err": (5, (N,), np.float64), "attempts": (9, (N, 64, 2), np.float64),
                "mult": (7, (N, self.info.n_cls), np.float32)}[what]
        out = np.empty(spec[1], spec[2])
        self._ck(self.L.epi_get_state(self._h, spec[0], out.ctypes.data, out.nbytes))
        return out

    def plan_tables(self):
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

    def attach_spec(self):
        """Compile the scenario-specialised build of this engine's kernel FOR THE HOST (the generated header +
        csrc/epi_spec_main.cu with -DEPI_HOST_EMU) and attach it: the generated per-cell program, prologue tables and
        constexpr layout are then what the emulated warps run."""
        import re
        import subprocess
        need = ctypes.c_size_t()
        self._ck(self.L.epi_spec_source(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        self._ck(self.L.epi_spec_source(self._h, buf, need.value, None))
        text = buf.value.decode()
        key = re.search(r'EPI_SPEC_KEY "([0-9a-f]+)"', text).group(1)
        csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "epipolicy_rl_b200", "csrc")
        d = os.path.join(HERE, "_spec")
        os.makedirs(d, exist_ok=True)
        hdr, out = os.path.join(d, f"spec_{key}.h"), os.path.join(d, f"spec_{key}.so")
        newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc))
        newest = max(newest, os.path.getmtime(os.path.join(HERE, "cuda_runtime.h")))
        if not os.path.exists(out) or os.path.getmtime(out) < newest:
            with open(hdr, "w") as f:
                f.write(text)
            subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                                   "-DEPI_HOST_EMU", f'-DEPI_SPEC_HEADER="{hdr}"', "-I", HERE, "-I", csrc,
                                   "-include", os.path.join(HERE, "cuda_runtime.h"),  # nvcc includes it implicitly
                                   "-Wno-unused-variable", "-pthread", "-o", out + ".tmp", os.path.join(csrc, "epi_spec_main.cu")])
            os.replace(out + ".tmp", out)
        self._ck(self.L.epi_spec_attach(self._h, out.encode()))
        self.L.epi_info(self._h, ctypes.byref(self.info))
        return key
