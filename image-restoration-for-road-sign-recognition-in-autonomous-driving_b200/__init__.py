"""B200-native (sm_100a) degrade -> restore -> classify path behind the reference's PyTorch module contract.

Importable as `b200restore` (see b200restore.py at the repo root; this directory's name is not a Python identifier).
"""
from . import _lib, build, ops, packing  # noqa: F401
from ._lib import B2RError  # noqa: F401

__all__ = ["_lib", "build", "ops", "packing", "B2RError"]
