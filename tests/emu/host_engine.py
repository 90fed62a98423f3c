"""The PRODUCT's host class (epipolicy_rl_b200.engine.BatchedEpidemic: step / step_plan / run / state_dev / plan_tables as
shipped) driven against the HOST EMULATION build of the library (tests only): torch CPU tensors stand in for device
tensors - the emulation's "device pointers" are host pointers - so the Python side of every C-ABI call is executed on
the CPU exactly as it is on the GPU box."""
import ctypes

import numpy as np
import torch

from emu_engine import emu_lib

from epipolicy_rl_b200 import lib as _lib
from epipolicy_rl_b200.engine import BatchedEpidemic


class HostBatchedEpidemic(BatchedEpidemic):
    def __init__(self, desc, n_envs, precision="fp64", spec=False):  # noqa: super().__init__ needs a CUDA device
        self.L = emu_lib()
        self.device = torch.device("cpu")
        self.desc, self.n_envs = desc, int(n_envs)
        self.precision = {"fp64": _lib.EPI_FP64, "fp32": _lib.EPI_FP32}[precision]
        self.dtype = torch.float64 if precision == "fp64" else torch.float32
        self.np_dtype = np.float64 if precision == "fp64" else np.float32
        self._c = desc.as_c()
        h = ctypes.c_void_p()
        rc = self.L.epi_create(ctypes.byref(self._c), self.n_envs, 0, self.precision, ctypes.byref(h))
        if rc != 0:
            raise _lib.EpiError(rc, self.L.epi_last_error(None).decode())
        self._h = h
        self.info = _lib.EpiInfo()
        self._ck(self.L.epi_info(self._h, ctypes.byref(self.info)))
        self.obs_dim, self.action_dim = self.info.obs_dim, self.info.action_dim
        self.obs = torch.empty((self.n_envs, self.obs_dim), dtype=self.dtype)
        self.reward = torch.zeros(self.n_envs, dtype=torch.float64)
        self.done = torch.zeros(self.n_envs, dtype=torch.uint8)
        self.reset()

    def _stream(self):
        return None
