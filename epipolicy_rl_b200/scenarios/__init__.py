"""Scenario descriptors shipped with the package: the reference's `Static` (obj/static.py) of every shipped JSON
scenario (jsons/*.json) and of the synthetic COVID-like scenarios (synth.py), exported once from the unmodified
reference by tools/gen_golden.py / tools/gen_syn.py (no parser, no numba JIT at start-up)."""
from __future__ import annotations

import glob
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def path(name: str) -> str:
    p = os.path.join(_HERE, f"desc_{name}.npz")
    if not os.path.exists(p):
        raise FileNotFoundError(f"no shipped scenario '{name}' ({p}); export one with tools/gen_golden.py or "
                                "tools/gen_syn.py (needs the reference importable), or build an EpiDesc with export.py")
    return p


def names() -> list[str]:
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(_HERE, "desc_*.npz")))


def load(name: str):
    from ..desc import EpiDesc
    return EpiDesc.load(path(name))
