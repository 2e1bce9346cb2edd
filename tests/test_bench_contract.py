"""bench.py's reference arm runs on the CPU: check the JSON line it prints against the contract (keys the driver reads)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    assert [l for l in r.stdout.splitlines() if l.strip()] == lines, "stdout must carry the JSON line and nothing else"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec restore->VGG16 classify" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "degrade16_resunet_vgg16_top1"
    cb = d["cpu_baseline"]
    have_ref = (ROOT / "baseline" / "_ref" / "MANIFEST.json").exists()
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["images_per_gpu_per_step"] == 4096 and d["config"]["sample_images_per_step"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
