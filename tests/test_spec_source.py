"""Scenario-specialised code generation (epi_spec_source_desc, include/epi_b200.h) needs no GPU: the generated header
is deterministic, keyed, covers every edge / compartment / accumulator of the scenario, and compiles with nvcc."""
import os
import re
import subprocess

import pytest

from conftest import SCENARIOS
from helpers import load

from epipolicy_rl_b200 import spec


@pytest.mark.parametrize("name", SCENARIOS + ["syn_s"])
def test_generated_header_covers_the_scenario(name):
    d, _ = load(name)
    t64, k64 = spec.source(d, 0)
    t32, k32 = spec.source(d, 1)
    assert k64 != k32 and spec.source(d, 1) == (t32, k32)  # keyed by precision, deterministic
    assert "typedef double real;" in t64 and "typedef float real;" in t32
    dims = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b([A-Za-z_]+) = (-?\d+)[,;]", t32))
    assert (dims["C"], dims["L"], dims["G"], dims["F"], dims["E"]) == (d.C, d.L, d.G, d.F, d.E)
    body = t32[t32.index("void flows("):t32.index("void dx(")]
    assert len(re.findall(r"\bf\[\d+\] = ", body)) == d.E          # every regular edge
    dx = t32[t32.index("void dx("):t32.index("void acc(")]
    assert len(re.findall(r"^\s+d\[\d+\] = 0;", dx, re.M)) == d.C   # every compartment
    assert ("EPI_SPEC_BIG 1" in t32) == (d.L * d.G > 32)


def test_stale_cache_entries_are_not_reused():
    d, _ = load("SIR_A")
    _, key = spec.source(d, 1)
    assert spec._src_hash() in os.path.basename(spec.so_path(key))  # kernel sources are part of the file name


@pytest.mark.skipif(spec.nvcc_path() is None, reason="nvcc not available")
def test_generated_header_compiles_for_sm100a(tmp_path):
    d, _ = load("SIRV_A")
    text, key = spec.source(d, 0)
    hdr = tmp_path / "spec.h"
    hdr.write_text(text)
    out = tmp_path / "spec.so"
    subprocess.check_call([spec.nvcc_path(), *spec.NVCC_FLAGS, f'-DEPI_SPEC_HEADER="{hdr}"', "-I", spec.CSRC, "-o", str(out),
                           os.path.join(spec.CSRC, "epi_spec_main.cu")])
    syms = subprocess.check_output(["nm", "-D", str(out)], text=True)
    assert "epi_spec_launch" in syms and "epi_spec_key" in syms
