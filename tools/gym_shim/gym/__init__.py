"""Minimal stand-in for the `gym` package (absent in this image) so that the
reference's epipolicy_environment.py imports. Only what that file uses."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def __init__(self):
        pass
