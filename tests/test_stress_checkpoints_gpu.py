"""bf16 error under a hostile checkpoint (VERDICT r1 'weak' #2): the tolerances in tests/test_models_gpu.py were measured on
variance-preserving synthetic weights.  Here the restorers run `synth.stress_state_dict` — un-damped residual gammas,
BatchNorm running_var log-uniform in [1e-3, 10], PReLU slopes in [0, 1], biases N(0, 0.5) — whose activations grow by
orders of magnitude through the network, and the error is stated RELATIVE to the reference output's own range:

    rel_max = max |ours - ref| / (max ref - min ref)        psnr_range = 20 log10(range / rmse)

The `final` 1x1 conv of the hostile checkpoint is rescaled (with the fp32 oracle, on a calibration batch) so that the
module output is centred in [0, 1] like a trained restorer's: the clamp / u8 comparison then sees mid-range values instead
of an all-saturated image.  What survives: rel_max <= 2e-2, psnr_range >= 45 dB, restored bytes mean |diff| < 1 LSB,
BN folding in fp64 keeps a 30x per-channel scale exact (the folded weights are rounded to bf16 once, like any weight)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _calibrated(arch, seed, x):
    """stress checkpoint whose `final` layer maps the oracle's output on x to mean 0.5, std 0.2"""
    from b200restore import synth
    from oracle import models_oracle as MO
    sd = synth.stress_state_dict(arch, seed)
    fn = MO.simple_unet_forward if arch == "simple_unet" else MO.resunet_forward
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        y = fn(sdc, x)
    mu = y.mean(dim=(0, 2, 3))
    sg = y.std(dim=(0, 2, 3))
    a = (0.2 / sg).cpu()
    sd["final.weight"] = sd["final.weight"] * a.view(3, 1, 1, 1)
    sd["final.bias"] = (sd["final.bias"] - mu.cpu()) * a + 0.5
    return sd, float(y.abs().max())


@pytest.mark.parametrize("arch,seed", [("resunet", 1), ("resunet", 2), ("simple_unet", 3)])
def test_stress_checkpoint_relative_error(arch, seed):
    from b200restore import models, synth
    from oracle import models_oracle as MO
    n, hw = 8, 224
    imgs, _ = synth.indexed_images(0, n, hw, hw, seed=21)
    x = MO.to_tensor_u8(imgs.cuda())
    sd, raw_max = _calibrated(arch, seed, x)
    fn = MO.simple_unet_forward if arch == "simple_unet" else MO.resunet_forward
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = fn(sdc, x)
    m = (models.SimpleUNet if arch == "simple_unet" else models.ResUNet)()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    out = m(x)
    rng_ = float(ref.max() - ref.min())
    err = (out - ref).abs()
    rel_max = float(err.max()) / rng_
    rmse = float(((out - ref) ** 2).mean().sqrt())
    psnr = 20 * np.log10(rng_ / rmse)
    u8, u8_ref = m.restore_u8(imgs.cuda()), MO.quantize_restored(ref)
    d = (u8.int() - u8_ref.int()).abs().float()
    print(f"\n[stress {arch} seed {seed}] raw |out| max before calibration {raw_max:.3g}; reference range {rng_:.3f}; "
          f"max |err| {float(err.max()):.3g} ({100 * rel_max:.2f} % of range), PSNR vs range {psnr:.1f} dB; restored bytes "
          f"mean |diff| {float(d.mean()):.3f} LSB, max {int(d.max())}, frac > 2 LSB {float((d > 2).float().mean()):.2e}")
    assert bool(torch.isfinite(out).all())
    assert rel_max <= 2e-2 and psnr >= 45.0
    assert float(d.mean()) < 1.0 and float((d > 4).float().mean()) < 1e-3
