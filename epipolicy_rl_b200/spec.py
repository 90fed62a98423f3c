"""Scenario-specialised builds of the cell kernel.

The reference JIT-compiles its right-hand side with numba (105 s on first use); the
analogue here: libepi_b200 generates the scenario's per-cell program as straight-line CUDA
(`epi_spec_source_desc`, include/epi_b200.h), this module compiles it once with nvcc for sm_100a into
`epipolicy_rl_b200/_spec/spec_<key>.so` (in-tree, so prebuilt files travel with the repo) and the engine
attaches it (`epi_spec_attach`). Without nvcc or with EPI_SPEC=0 the interpreted cell kernel runs — still
CUDA, never a CPU path."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

from . import lib as _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
SPEC_DIR = os.path.join(_HERE, "_spec")
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def source(desc, precision: int, L=None) -> tuple[str, str]:
    """(generated header text, key) for a descriptor; needs no GPU."""
    L = L or _lib.load()
    c = desc.as_c()
    need = ctypes.c_size_t(0)
    rc = L.epi_spec_source_desc(ctypes.byref(c), precision, None, 0, ctypes.byref(need))
    if rc != 0:
        raise _lib.EpiError(rc, L.epi_last_error(None).decode())
    buf = ctypes.create_string_buffer(need.value)
    rc = L.epi_spec_source_desc(ctypes.byref(c), precision, buf, need.value, ctypes.byref(need))
    if rc != 0:
        raise _lib.EpiError(rc, L.epi_last_error(None).decode())
    text = buf.value.decode()
    key = re.search(r'EPI_SPEC_KEY "([0-9a-f]{16})"', text).group(1)
    return text, key


_SRC_HASH = None


def _extra_flags():
    """tuning aids forwarded to the specialised build (part of the cache key)"""
    mb = os.environ.get("EPI_ENV_MINB")
    return [f"-DEPI_ENV_MINB={int(mb)}"] if mb else []


def _src_hash() -> str:
    """hash of the kernel sources a specialised build compiles: a stale cache entry is never reused"""
    global _SRC_HASH
    if _SRC_HASH is None:
        import hashlib
        h = hashlib.sha1()
        for f in ("epi_cell.cuh", "epi_env.cuh", "epi_grp.cuh", "epi_rk.cuh", "epi_fdiv.cuh", "epi_kernels.cuh", "epi_plan.h", "epi_spec_main.cu"):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
        h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
        _SRC_HASH = h.hexdigest()[:10]
    return _SRC_HASH


def so_path(key: str) -> str:
    return os.path.join(SPEC_DIR, f"spec_{key}_{_src_hash()}.so")


def nvcc_path():
    p = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return p if os.path.exists(p) else None


def compile_spec(text: str, key: str) -> str | None:
    """nvcc the generated header + csrc/epi_spec_main.cu -> _spec/spec_<key>.so (None if nvcc is missing)."""
    out = so_path(key)
    if os.path.exists(out):
        return out
    nvcc = nvcc_path()
    if nvcc is None:
        return None
    os.makedirs(SPEC_DIR, exist_ok=True)
    hdr = out[:-3] + ".h"
    with open(hdr + f".{os.getpid()}", "w") as f:
        f.write(text)
    os.replace(hdr + f".{os.getpid()}", hdr)
    tmp = f"{out}.{os.getpid()}.tmp"
    subprocess.check_call([nvcc, *NVCC_FLAGS, *_extra_flags(), f'-DEPI_SPEC_HEADER="{hdr}"', "-I", CSRC, "-o", tmp,
                           os.path.join(CSRC, "epi_spec_main.cu")])
    os.replace(tmp, out)
    return out


def ensure(desc, precision: int, compile_missing: bool = True) -> str | None:
    """Path of the specialised library for (desc, precision), building it if needed and possible."""
    if os.environ.get("EPI_SPEC", "1") == "0":
        return None
    try:
        text, key = source(desc, precision)
    except _lib.EpiError as e:
        if e.code != -2:  # EPI_E_UNSUPPORTED: the scenario runs on the team kernels, which have no specialised build
            raise
        return None
    out = so_path(key)
    if os.path.exists(out):
        return out
    if not compile_missing:
        return None
    try:
        return compile_spec(text, key)
    except (OSError, subprocess.CalledProcessError) as e:
        # unwritable _spec directory, nvcc failure on the generated header, ...: the interpreted kernel (still CUDA) runs
        import sys
        print(f"[epi-b200] specialised build failed ({type(e).__name__}: {e}); using the interpreted kernel", file=sys.stderr)
        return None
