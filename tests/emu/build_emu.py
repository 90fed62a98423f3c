"""Builds tests/emu/libepi_b200_emu.so: the product's csrc compiled for the HOST (see cuda_runtime.h here)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build():
    out = os.path.join(HERE, "libepi_b200_emu.so")
    csrc = os.path.join(ROOT, "epipolicy_rl_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in ("epi_engine.cu", "epi_kernels.cuh", "epi_cell.cuh", "epi_env.cuh", "epi_grp.cuh", "epi_rk.cuh", "epi_fdiv.cuh", "epi_plan.h")] + [
        os.path.join(HERE, "cuda_runtime.h"), os.path.join(ROOT, "include", "epi_b200.h"), os.path.join(ROOT, "include", "epi_desc.h")]
    if (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-DEPI_HOST_EMU", "-I", HERE, "-Wno-unused-variable", "-pthread", "-o", out, srcs[0]])
    return out


if __name__ == "__main__":
    print(build())
