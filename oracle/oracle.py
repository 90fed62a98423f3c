"""ctypes wrapper of oracle/libepi_oracle.so — TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from epipolicy_rl_b200.desc import CEpiDesc, EpiDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libepi_oracle.so")
    src = os.path.join(_HERE, "epi_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "epi_desc.h")
    stale = (not os.path.exists(so)) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libepi_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        P = ctypes.c_void_p
        L.oracle_create.restype = P
        L.oracle_create.argtypes = [ctypes.POINTER(CEpiDesc)]
        L.oracle_reset.argtypes = [P]
        L.oracle_destroy.argtypes = [P]
        L.oracle_day_step.restype = ctypes.c_double
        L.oracle_day_step.argtypes = [P, P, P, P]
        L.oracle_env_step.restype = ctypes.c_double
        L.oracle_env_step.argtypes = [P, P, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.oracle_get_X.argtypes = [P, P]
        L.oracle_get_flat.argtypes = [P, P]
        L.oracle_set_flat.argtypes = [P, P]
        L.oracle_get_cost.argtypes = [P, P]
        L.oracle_n_flat.argtypes = [P]
        L.oracle_day.argtypes = [P]
        L.oracle_get_trace.argtypes = [P, P, P]
        L.oracle_get_param.argtypes = [P, ctypes.c_int, P, P, P, P, P]
        L.oracle_get_attempts.argtypes = [P, P, ctypes.c_int]
        L.oracle_get_moves.argtypes = [P, ctypes.c_int, P, ctypes.c_int]
        L.oracle_rollout.restype = ctypes.c_double
        L.oracle_rollout.argtypes = [ctypes.POINTER(CEpiDesc), P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """One environment of the restated reference simulator."""

    def __init__(self, desc: EpiDesc):
        self.desc = desc
        self._c = desc.as_c()
        self._h = lib().oracle_create(ctypes.byref(self._c))
        self.reset()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def reset(self):
        lib().oracle_reset(self._h)
        return self.X.reshape(-1)

    @property
    def X(self):
        d = self.desc
        out = np.empty((d.C, d.L, d.G))
        lib().oracle_get_X(self._h, _p(out))
        return out

    @property
    def cost(self):
        d = self.desc
        out = np.empty((d.I, d.L))
        lib().oracle_get_cost(self._h, _p(out))
        return out

    @property
    def flat(self):
        out = np.empty(lib().oracle_n_flat(self._h))
        lib().oracle_get_flat(self._h, _p(out))
        return out

    @property
    def trace(self):
        t = np.zeros(4, np.int32)
        e = ctypes.c_double()
        lib().oracle_get_trace(self._h, _p(t), ctypes.byref(e))
        return dict(nfev=int(t[0]), nsteps=int(t[1]), nrej=int(t[2]), status=int(t[3]), last_err=e.value)

    @property
    def attempts(self):
        """(h, error_norm) of every attempted RK45 step of the last simulated day"""
        a = np.zeros((64, 2))
        n = lib().oracle_get_attempts(self._h, _p(a), 64)
        return a[:n]

    def params(self, which=0):
        d = self.desc
        mp = np.empty((d.P, d.L, d.F, d.G), np.float32)
        ft = np.empty((d.L, d.F, d.G), np.float32)
        fi = np.empty((d.L, d.F, d.G, d.G), np.float32)
        mob = np.empty(d.G * d.MODE_NNZ_TOTAL, np.float32)
        cost = np.empty((d.I, d.L))
        lib().oracle_get_param(self._h, which, _p(mp), _p(ft), _p(fi), _p(mob), _p(cost))
        n = lib().oracle_get_moves(self._h, which, None, 0)
        mv = np.zeros((max(n, 1), 6))
        lib().oracle_get_moves(self._h, which, _p(mv), n)
        return dict(mp=mp, ft=ft, fi=fi, mob=mob, cost=cost, moves=mv[:n])

    def day_step(self, cpv, active, init_programs=False):
        cpv = np.ascontiguousarray(cpv, np.float32)
        active = np.ascontiguousarray(active, np.int32)
        prog = _p(self.desc.prog_init) if init_programs else None
        return lib().oracle_day_step(self._h, _p(cpv), _p(active), prog)

    def step(self, action, n_days=7):
        """EpiEnv.step semantics: returns (obs, reward, done)."""
        a = np.ascontiguousarray(action, np.float32)
        done = ctypes.c_int()
        r = lib().oracle_env_step(self._h, _p(a), n_days, ctypes.byref(done))
        return self.X.reshape(-1), r, bool(done.value)


def rollout(desc: EpiDesc, actions: np.ndarray, n_days: int = 7, want_final=True):
    """actions [n_envs, n_steps, A] -> (sum of rewards, final X [n_envs, C*L*G])."""
    actions = np.ascontiguousarray(actions, np.float32)
    n_envs, n_steps = actions.shape[:2]
    fin = np.empty((n_envs, desc.C * desc.L * desc.G)) if want_final else None
    c = desc.as_c()
    s = lib().oracle_rollout(ctypes.byref(c), _p(actions), n_envs, n_steps, n_days, _p(fin) if want_final else None)
    return s, fin
