"""Regenerate include/epi_desc.h from the field list in epipolicy_rl_b200/desc.py."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from epipolicy_rl_b200.desc import c_header_text  # noqa: E402

path = os.path.join(os.path.dirname(__file__), "..", "include", "epi_desc.h")
open(path, "w").write(c_header_text())
print("wrote", os.path.normpath(path))
