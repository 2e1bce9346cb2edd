#!/usr/bin/env python
"""Product path only (no oracle, no torch reference) on small inputs, for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck --error-exitcode 9 python tools/sanitize.py

degrade (script-16 recipe, Philox noise) -> ResUNet -> u8 -> VGG16 -> top-1 count at 64 x 64 (8 images) and 224 x 224
(2 images), SimpleUNet forward at 32 x 32, and the kernels named at the end (every tcgen05 conv kernel class must have run)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import b200restore as B  # noqa: E402
from b200restore import degrade, models, ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
r, j, u = models.ResUNet(), models.VGG16Judge(), models.SimpleUNet()
r.load_state_dict(synth.synthetic_state_dict("resunet", 31))
j.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
u.load_state_dict(synth.synthetic_state_dict("simple_unet", 33))
r, j, u = r.to(dev).eval(), j.to(dev).eval(), u.to(dev).eval()
timer = ops.KernelTimer()
with torch.no_grad(), ops.timing(timer):
    for n, hw in ((8, 64), (2, 224)):
        pipe = B.RestoreClassifyPipeline(r, j, micro_batch=n)
        imgs, labels = synth.indexed_images(0, n, hw, hw, seed=5)
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        pipe.run_micro_batch(imgs.to(dev), labels.to(dev), degrade.compound_params(n).to(dev), 2, 0, counts)
        torch.cuda.synchronize()
        assert int(counts[1]) == n
    y = u(torch.rand((2, 3, 32, 32), device=dev))
    torch.cuda.synchronize()
    assert y.shape == (2, 3, 32, 32) and bool(torch.isfinite(y).all())
print("sanitize run ok; conv kernels seen:", sorted({rec[4] for rec in timer.records}))
