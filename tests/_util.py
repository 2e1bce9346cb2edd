"""Shared helpers for the test-suite (test infrastructure only)."""
from __future__ import annotations

import importlib.util
import io
import sys
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
REF = Path("/root/reference")
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def have_reference() -> bool:
    return (REF / "14_train_unified_advanced.py").exists()


def load_ref(filename: str):
    """Import one of the reference's numbered scripts (module-level code only defines things; 07 prints)."""
    spec = importlib.util.spec_from_file_location("ref_" + filename.split("_")[0], REF / filename)
    mod = importlib.util.module_from_spec(spec)
    with redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


class ReplayNormal:
    def __init__(self, z):
        self.z = z

    def __call__(self, loc, scale, size):
        assert tuple(size) == self.z.shape
        return loc + scale * self.z


class NpShim:
    """numpy as seen by a reference module, with random.normal replaced; numpy itself stays untouched."""

    def __init__(self, normal):
        self.random = types.SimpleNamespace(normal=normal)

    def __getattr__(self, name):
        return getattr(np, name)


class ReplayRandom:
    def __init__(self, draws):
        self.draws = list(draws)

    def random(self):
        return self.draws.pop(0)

    def uniform(self, a, b):
        return self.draws.pop(0)

    def randint(self, a, b):
        return self.draws.pop(0)


def golden(name: str):
    return np.load(GOLDEN / name, allow_pickle=False)
