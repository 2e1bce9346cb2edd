#!/usr/bin/env python
"""Streaming kernels of csrc/generators.cu against the HBM roofline: algorithmic bytes (read + written) per image /
CUDA-event time, batch 1024 of 224x224x3 (154 MB per tensor: larger than L2).

    python tools/generators_bench.py [--json out.json]
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from b200restore import generators as G, ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    n, hw = 1024, 224
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = float(peaks.get("hbm_gbps", 6534.5))
    img = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device=dev)
    other = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device=dev)
    out = torch.empty_like(img)
    lut = torch.randint(0, 256, (n, 256), dtype=torch.uint8, device=dev)
    sigma = torch.full((n,), 0.02 ** 0.5, dtype=torch.float32, device=dev)
    mm = ops.minmax_u8(img)
    px = hw * hw * 3
    cases = [
        ("lut_u8 (fog 04 / 13)", lambda: ops.lut_u8(img, lut, out=out), 2 * px),
        ("minmax_u8", lambda: ops.minmax_u8(img), px),
        ("normalize_minmax_u8 (03 stretch)", lambda: ops.normalize_minmax_u8(img, mm, out=out), 2 * px),
        ("noise02 script-02 rule, Philox (two passes)", lambda: ops.noise02(img, sigma, seed=1, out=out), 2 * px),
        ("noise02 unit clip (13), Philox", lambda: ops.noise02(img, sigma, seed=1, out=out, clip_rule=1), 2 * px),
        ("sse_u8 (PSNR)", lambda: ops.sse_u8(img, other), 2 * px),
        ("ssim_u8 (08:123, 7x7 windows, float64)", lambda: ops.ssim_u8(img, other), 2 * px),
        ("apply_motion_blur d=12 + stretch (03)", lambda: G.apply_motion_blur(img, 12, 45), 2 * px),
    ]
    res = []
    for name, fn, bytes_per_img in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbps = n * bytes_per_img / (ms * 1e-3) / 1e9
        print(f"{name:46s} {ms * 1e3 / n:7.3f} us/img  {gbps:8.1f} GB/s algorithmic = {100 * gbps / hbm:5.1f}% of {hbm:.0f} GB/s")
        res.append({"kernel": name, "us_per_image": ms * 1e3 / n, "algorithmic_gbps": gbps, "frac_hbm": gbps / hbm})
    # ragged batch of GTSRB-like sizes (15..250 px) -> 224 x 224: kernel only (tables and packing prepared once)
    from b200restore import imageio as IO
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (int(rng.integers(15, 251)), int(rng.integers(15, 251)), 3), dtype=np.uint8) for _ in range(n)]
    rout, plan = IO.resize_batch(imgs, _return_plan=True)
    in_bytes = sum(im.size for im in imgs)
    for _ in range(3):
        ops.resize_bilinear_u8(out=rout, **plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.resize_bilinear_u8(out=rout, **plan)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbps = (in_bytes + rout.numel()) / (ms * 1e-3) / 1e9
    name = "resize_bilinear_u8 (Pillow BILINEAR, 15..250 px -> 224)"
    print(f"{name:46s} {ms * 1e3 / n:7.3f} us/img  {gbps:8.1f} GB/s algorithmic = {100 * gbps / hbm:5.1f}% of {hbm:.0f} GB/s")
    res.append({"kernel": name, "us_per_image": ms * 1e3 / n, "algorithmic_gbps": gbps, "frac_hbm": gbps / hbm})
    cout, cplan = IO.resize_batch_cv(imgs, _return_plan=True)
    for _ in range(3):
        ops.resize_cv_linear_u8(out=cout, **cplan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.resize_cv_linear_u8(out=cout, **cplan)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbps = (in_bytes + cout.numel()) / (ms * 1e-3) / 1e9
    name = "resize_cv_linear_u8 (cv2.resize, 15..250 px -> 224)"
    print(f"{name:46s} {ms * 1e3 / n:7.3f} us/img  {gbps:8.1f} GB/s algorithmic = {100 * gbps / hbm:5.1f}% of {hbm:.0f} GB/s")
    res.append({"kernel": name, "us_per_image": ms * 1e3 / n, "algorithmic_gbps": gbps, "frac_hbm": gbps / hbm})
    if "--json" in sys.argv:
        Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps({"batch": n, "hw": hw, "hbm_gbps": hbm, "kernels": res}, indent=1))


if __name__ == "__main__":
    main()
