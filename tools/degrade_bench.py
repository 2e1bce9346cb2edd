#!/usr/bin/env python
"""Time b2r_degrade at 224x224, batch 1024 (154 MB in + 154 MB out >> L2) for the recipes the path uses; prints the
algorithmic HBM rate (6 B/pixel: u8 in + u8 out) against MEASURED_PEAKS.json:hbm_gbs."""
import json
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from b200restore import degrade as D

n, hw = 1024, 224
dev = torch.device("cuda", 0)
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6534.5) if (ROOT / "MEASURED_PEAKS.json").exists() else 6534.5
img = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device=dev)
out = torch.empty_like(img)
res = {}
recipes = {
    "script16 blur(10,45)+fog+noise": D.compound_params(n),
    "script14 random mix (fog/noise/blur p=.5)": D.random_params(n, np.random.default_rng(0)),
    "script15 fog+noise+clip -> blur": D.demo_params(n),
    "fog+noise only": (lambda p: ([p.set_fog(i, 0.5) or p.set_noise(i, 0.02) for i in range(n)], p)[1])(D.DegradeParams(n)),
    "fog only": D.fog_params(n, np.random.default_rng(1)),
    "blur(10,45) only": D.blur_params(n, 10, 45),
}
for name, p in recipes.items():
    dp = p.to(dev)
    for _ in range(3):
        D.degrade(img, dp, seed=1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        D.degrade(img, dp, seed=1, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = n * hw * hw * 6 / ms / 1e6
    res[name] = {"ms": ms, "us_per_image": ms * 1e3 / n, "GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    print(f"{name:44s} {ms * 1e3 / n:6.3f} us/img  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of measured HBM peak")
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(res, indent=1))
