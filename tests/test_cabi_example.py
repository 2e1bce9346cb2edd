"""examples/cabi_degrade_metrics.c: the C-ABI used from plain C (dlopen + CUDA runtime allocations, no PyTorch).
CPU: the example compiles against include/b2r.h.  GPU: it runs and its own checks pass (fog bytes against the
reference arithmetic restated in C, exact SSE, SSIM range, error convention)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "examples" / "cabi_degrade_metrics.c"


def _compile(out: Path):
    cuda = Path("/usr/local/cuda")
    if shutil.which("gcc") is None or not (cuda / "include" / "cuda_runtime.h").exists():
        pytest.skip("gcc / CUDA toolkit headers not available")
    cmd = ["gcc", "-O2", "-Wall", "-I", str(ROOT / "include"), "-I", str(cuda / "include"), str(SRC), "-o", str(out),
           "-L", str(cuda / "lib64"), "-lcudart", "-ldl", "-lm", f"-Wl,-rpath,{cuda / 'lib64'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_example_compiles_against_the_header(tmp_path):
    exe = _compile(tmp_path / "cabi_demo")
    assert exe.exists()


@pytest.mark.gpu
def test_example_runs_without_pytorch(tmp_path):
    from b200restore import build
    exe = _compile(tmp_path / "cabi_demo")
    r = subprocess.run([str(exe), str(build.LIB_PATH)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok") and "N = 0 rejected: -22" in r.stdout
