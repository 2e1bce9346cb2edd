#!/usr/bin/env python
"""BASELINE.json configs[4]: resolution x batch x GPU-count sweep of the headline workload, with the reference's CPU path per
resolution beside it.  One compact JSON line per point; redirect into profiles/.

    python tools/sweep.py [--hw 64,128,224,256] [--batch 256,1024,4096,16384] [--gpus 1] [--cpu] > profiles/r02_sweep.jsonl

--gpus takes a list (e.g. 1,2,4,8): N > 1 runs bench.py under torchrun with one rank per GPU (weak scaling: --batch is per
GPU).  --cpu adds one `bench.py --impl reference` run per resolution (the unmodified reference scripts on the host cores,
bounded sample; its images/s does not depend on the batch or the GPU count)."""
import argparse
import json
import socket
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def last_json(text):
    for ln in reversed(text.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise ValueError("no JSON line")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", default="64,128,224,256")
    ap.add_argument("--batch", default="256,1024,4096,16384")
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    ints = lambda s: [int(v) for v in s.split(",") if v]  # noqa: E731
    for hw in ints(a.hw):
        cpu = None
        if a.cpu:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--hw", str(hw), "--steps", "1",
                                "--warmup", "1", "--cpu-sample", str(64 if hw >= 224 else 128)], capture_output=True, text=True)
            try:
                d = last_json(r.stdout)
                cpu = {"images_per_s": d["value"], "cores": d["cpu_baseline"]["cores"], "kind": d["cpu_baseline"]["kind"]}
            except Exception as e:  # noqa: BLE001
                cpu = {"error": str(e), "stderr": r.stderr[-300:]}
        for gpus in ints(a.gpus):
            for batch in ints(a.batch):
                args = [str(ROOT / "bench.py"), "--gpus", str(gpus), "--hw", str(hw), "--batch", str(batch), "--steps", str(a.steps),
                        "--warmup", "3", "--no-cpu-baseline", "--no-comparator"]
                cmd = ([sys.executable] + args) if gpus == 1 else \
                    [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr",
                     "127.0.0.1", "--master-port", str(free_port())] + args
                r = subprocess.run(cmd, capture_output=True, text=True)
                try:
                    d = last_json(r.stdout)
                    print(json.dumps({"hw": hw, "gpus": gpus, "batch_per_gpu": batch, "micro_batch": d["config"]["micro_batch"],
                                      "images_per_s": d["value"], "e2e": d["e2e"]["value"],
                                      "conv_tflops": d["roofline"]["achieved"], "conv_frac_sustained": d["roofline"]["frac"],
                                      "sm_mhz": d["clocks"]["sm_mhz"], "per_rank_ms": d["per_rank"]["ms_per_step"],
                                      "cpu_reference": cpu, "speedup_vs_cpu": (d["e2e"]["value"] / cpu["images_per_s"])
                                      if cpu and "images_per_s" in cpu else None}), flush=True)
                except Exception as e:  # noqa: BLE001
                    print(json.dumps({"hw": hw, "gpus": gpus, "batch_per_gpu": batch, "error": str(e),
                                      "stderr": r.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
