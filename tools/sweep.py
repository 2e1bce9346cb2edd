#!/usr/bin/env python
"""BASELINE.json configs[4]: resolution x batch sweep of the headline workload on one GPU (bench.py without the CPU leg).
Prints one compact JSON line per point; redirect into profiles/.

    python tools/sweep.py > profiles/r01_sweep_resolution_batch_v9.jsonl
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for hw in (64, 128, 224, 256):
    for batch in (1024, 4096):
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--hw", str(hw), "--batch", str(batch), "--steps", "3",
                            "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(json.dumps({"hw": hw, "batch": batch, "images_per_s": d["value"], "e2e": d["e2e"]["value"],
                              "conv_tflops": d["roofline"]["achieved"], "sm_mhz": d["clocks"]["sm_mhz"]}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"hw": hw, "batch": batch, "error": str(e), "stderr": r.stderr[-300:]}), flush=True)
