"""ctypes binding of libepi_b200.so (include/epi_b200.h). Fails loudly if the library is missing:
there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes
import os

from .desc import CEpiDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libepi_b200.so")

EPI_FP64, EPI_FP32 = 0, 1
EXPORTS = ["epi_create", "epi_destroy", "epi_last_error", "epi_info", "epi_reset", "epi_step", "epi_step_plan", "epi_run_plan",
           "epi_step_host", "epi_get_state", "epi_set_state", "epi_launch_count", "epi_spec_source", "epi_spec_source_desc",
           "epi_spec_attach", "epi_host_buffers", "epi_bind_mirror", "epi_get_state_dev", "epi_get_plan"]


class EpiInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n_envs", "precision", "device", "obs_dim", "action_dim", "ncp", "n_itv", "horizon", "n_flat", "n_acc",
        "n_slot", "n_lc", "n_groups_inf", "n_cls", "team_kind", "threads", "grid", "envs_per_cta", "specialized",
        "reserved")] + [
        (n, ctypes.c_int64) for n in ("smem_bytes", "small_bytes", "big_bytes", "scratch_bytes", "state_bytes_per_env", "mv_len")]


class EpiRunOut(ctypes.Structure):
    """include/epi_b200.h: per-day outputs of epi_run_plan (device pointers, 0 = not wanted)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("obs_days", "reward_days", "x_days", "acc_days", "mult_days")]


class EpiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libepi_b200 error {code}: {msg}")
        self.code = code


_LIB = None


def load():
    """Load the CUDA library built by __graft_entry__.build(); raise if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). The B200 engine has no CPU fallback.")
    _LIB = bind(ctypes.CDLL(LIB_PATH))
    return _LIB


def bind(L):
    """Declare the C-ABI signatures (include/epi_b200.h) on a loaded library."""
    P, I, SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    L.epi_create.restype = I
    L.epi_create.argtypes = [ctypes.POINTER(CEpiDesc), I, I, I, ctypes.POINTER(P)]
    L.epi_destroy.restype = None
    L.epi_destroy.argtypes = [P]
    L.epi_last_error.restype = ctypes.c_char_p
    L.epi_last_error.argtypes = [P]
    L.epi_info.restype = I
    L.epi_info.argtypes = [P, ctypes.POINTER(EpiInfo)]
    L.epi_reset.restype = I
    L.epi_reset.argtypes = [P, P, P, P]
    L.epi_step.restype = I
    L.epi_step.argtypes = [P, P, I, P, P, P, P]
    L.epi_step_plan.restype = I
    L.epi_step_plan.argtypes = [P, P, P, I, I, P, P, P, P]
    L.epi_run_plan.restype = I
    L.epi_run_plan.argtypes = [P, P, P, I, I, ctypes.POINTER(EpiRunOut), P, P, P, P]
    L.epi_step_host.restype = I
    L.epi_step_host.argtypes = [P, P, I, P, P, P]
    L.epi_get_state.restype = I
    L.epi_get_state.argtypes = [P, I, P, SZ]
    L.epi_set_state.restype = I
    L.epi_set_state.argtypes = [P, I, P, SZ]
    L.epi_launch_count.restype = ctypes.c_int64
    L.epi_launch_count.argtypes = [P]
    L.epi_spec_source.restype = I
    L.epi_spec_source.argtypes = [P, P, SZ, ctypes.POINTER(SZ)]
    L.epi_spec_source_desc.restype = I
    L.epi_spec_source_desc.argtypes = [ctypes.POINTER(CEpiDesc), I, P, SZ, ctypes.POINTER(SZ)]
    L.epi_host_buffers.restype = I
    L.epi_host_buffers.argtypes = [P, ctypes.POINTER(P), ctypes.POINTER(P), ctypes.POINTER(P), ctypes.POINTER(P)]
    L.epi_bind_mirror.restype = I
    L.epi_bind_mirror.argtypes = [P, P, P, P]
    L.epi_spec_attach.restype = I
    L.epi_spec_attach.argtypes = [P, ctypes.c_char_p]
    L.epi_get_state_dev.restype = I
    L.epi_get_state_dev.argtypes = [P, I, P, SZ, P]
    L.epi_get_plan.restype = I
    L.epi_get_plan.argtypes = [P, I, P, SZ]
    return L
