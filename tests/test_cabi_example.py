"""examples/cabi_degrade_metrics.c: the C-ABI used from plain C (dlopen + CUDA runtime allocations, no PyTorch).
CPU: the example compiles against include/b2r.h.  GPU: it runs and its own checks pass (fog bytes against the
reference arithmetic restated in C, exact SSE, SSIM range, error convention)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "examples" / "cabi_degrade_metrics.c"


def _compile(out: Path, src: Path = None):
    global SRC
    if src is not None:
        saved, SRC = SRC, src
        try:
            return _compile(out)
        finally:
            SRC = saved
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


def test_pipeline_example_compiles_against_the_header(tmp_path):
    assert _compile(tmp_path / "cabi_pipeline", ROOT / "examples" / "cabi_pipeline.c").exists()


@pytest.mark.gpu
def test_pipeline_example_matches_the_python_host(tmp_path):
    """degrade -> ResUNet -> u8 -> VGG16 -> top-1 -> count from plain C through the whole-network entry points
    (b2r_net_create / b2r_resunet_forward / b2r_vgg16_forward) == the same path through the nn.Module classes, byte for byte."""
    import sys
    import torch
    import b200restore as B
    from b200restore import build, degrade as D, models, synth
    exe = _compile(tmp_path / "cabi_pipeline", ROOT / "examples" / "cabi_pipeline.c")
    bundle = tmp_path / "bundle.bin"
    subprocess.run([sys.executable, str(ROOT / "examples" / "export_bundle.py"), str(bundle), "--n", "16", "--hw", "64"], check=True)
    r = subprocess.run([str(exe), str(build.LIB_PATH), str(bundle)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = dict(l.split(" fnv1a ") for l in r.stdout.splitlines() if " fnv1a " in l)
    counts = [int(v) for v in next(l for l in r.stdout.splitlines() if l.startswith("counts")).split()[1:]]
    assert "Missing key" in r.stdout and r.stdout.strip().endswith("ok")

    def fnv1a(b: bytes) -> str:
        h = 1469598103934665603
        for x in b:
            h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return f"{h:016x}"

    rm, jm = models.ResUNet(), models.VGG16Judge()
    rm.load_state_dict(synth.synthetic_state_dict("resunet", 31))
    jm.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
    pipe = B.RestoreClassifyPipeline(rm.cuda(), jm.cuda(), micro_batch=16)
    imgs, labels = synth.indexed_images(0, 16, 64, 64, seed=7)
    c = torch.zeros(2, dtype=torch.int64, device="cuda")
    pred, extra = pipe.run_micro_batch(imgs.cuda(), labels.cuda(), D.compound_params(16).to("cuda"), 2, 0, c, keep=True)
    assert out["restored"] == fnv1a(extra["restored"].cpu().numpy().tobytes())
    assert out["pred"] == fnv1a(pred.cpu().numpy().tobytes())
    assert counts == c.tolist()
