"""Host-side checks that need no GPU: the generated header is current, the C-ABI library loads and
exports every symbol include/epi_b200.h declares, EpiDesc round-trips."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from epipolicy_rl_b200 import desc as D
from helpers import load


def test_generated_header_is_current():
    assert open(os.path.join(ROOT, "include", "epi_desc.h")).read() == D.c_header_text()


def test_ctypes_struct_matches_header_field_order():
    txt = open(os.path.join(ROOT, "include", "epi_desc.h")).read()
    body = txt[txt.index("typedef struct EpiDesc {"):txt.index("} EpiDesc;")]
    names = re.findall(r"\b(\w+); /\*", body)
    assert names == [f[0] for f in D.CEpiDesc._fields_]


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from epipolicy_rl_b200 import lib
    L = lib.load()
    hdr = open(os.path.join(ROOT, "include", "epi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(epi_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(lib.EXPORTS)
    for name in declared:
        assert getattr(L, name) is not None
    # no GPU here: create must fail cleanly with an error code and a message, not crash
    d, _ = load("SIR_A")
    c = d.as_c()
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        rc = L.epi_create(ctypes.byref(c), 4, 0, 0, ctypes.byref(h))
        assert rc == -3 and h.value is None
        assert b"cuda" in L.epi_last_error(None)
    assert L.epi_create(None, 4, 0, 0, ctypes.byref(h)) == -1


def test_desc_roundtrip(tmp_path):
    d, _ = load("COVID_C")
    p = str(tmp_path / "d.npz")
    d.save(p)
    e = D.EpiDesc.load(p)
    assert e.scalars == d.scalars
    for n, _, _ in D.ARRAYS:
        assert np.array_equal(e.arrays[n], d.arrays[n])
    assert e.act_domain.shape == (14, 2) and e.A == 6 and e.n_flat == 2457


def test_product_does_not_import_oracle():
    pk = os.path.join(ROOT, "epipolicy_rl_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").lower() or f in (), (f,)


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from epipolicy_rl_b200.engine import BatchedEpidemic
    d, _ = load("SIR_A")
    with pytest.raises(RuntimeError):
        BatchedEpidemic(d, 4)


def test_compat_module_has_the_reference_module_surface():
    """epipolicy_rl_b200/compat/epipolicy_environment.py stands in for the reference module of the same name
    (epipolicy_environment.py:1-71): PERIOD and an EpiEnv class with the gym surface."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "epipolicy_rl_b200", "compat", "epipolicy_environment.py")
    spec = importlib.util.spec_from_file_location("epipolicy_environment_compat", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.PERIOD == 7
    for name in ("reset", "step", "render", "close"):
        assert callable(getattr(mod.EpiEnv, name))
