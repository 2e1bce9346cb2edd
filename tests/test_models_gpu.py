"""GPU parity of the drop-in modules against the fp32 oracle (oracle/models_oracle.py, itself pinned to the
reference classes) on identical inputs and identical seeded checkpoints.

Floating point, bf16 activations with fp32 accumulation vs an fp32 reference: the tolerance is stated per test as
max-abs / PSNR on the f32 output (BASELINE.json north_star: "fp32-accumulated restorations must match within a
stated max-abs/PSNR tolerance").  The comparator runs on the same GPU in fp32 with TF32 disabled.
"""
import io
import math

import numpy as np
import pytest
import torch

from _util import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def psnr(a, b, peak=1.0):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * math.log10(peak * peak / mse)


def _load(arch, seed):
    from b200restore import models, synth
    sd = synth.synthetic_state_dict(arch, seed)
    m = {"simple_unet": models.SimpleUNet, "resunet": models.ResUNet, "vgg16": models.VGG16Judge}[arch]()
    m.load_state_dict(sd)
    return m.cuda().eval(), {k: v.cuda() for k, v in sd.items()}


# observed on B200: PSNR 59-63 dB, max-abs 3e-3 .. 7e-3 on outputs spanning about [0, 1]
RESTORER_TOL = {"psnr_db": 50.0, "max_abs": 2e-2}


@pytest.mark.parametrize("arch,shape", [("simple_unet", (2, 3, 64, 96)), ("simple_unet", (3, 3, 224, 224)),
                                        ("resunet", (2, 3, 64, 96)), ("resunet", (3, 3, 224, 224))])
def test_restorer_forward_vs_oracle(arch, shape):
    from oracle import models_oracle as O
    m, sd = _load(arch, 21)
    fn = O.simple_unet_forward if arch == "simple_unet" else O.resunet_forward
    from b200restore import synth
    img, _ = synth.sign_like_images(shape[0], shape[2], shape[3], seed=1)
    x = O.to_tensor_u8(img).cuda()
    y = m(x)
    with torch.no_grad():
        ref = fn(sd, x)
    assert y.shape == ref.shape and y.dtype == torch.float32
    p, mx = psnr(y, ref), float((y - ref).abs().max())
    print(f"\n[{arch} {shape}] PSNR {p:.1f} dB, max-abs {mx:.4g}, ref range [{float(ref.min()):.3f}, {float(ref.max()):.3f}]")
    assert p >= RESTORER_TOL["psnr_db"] and mx <= RESTORER_TOL["max_abs"]


def test_restorer_matches_committed_reference_output():
    """Directly against the reference classes' outputs stored by tests/golden/make_golden.py (same seed 11)."""
    g = golden("models_ref.npz")
    for arch in ("simple_unet", "resunet"):
        m, _ = _load(arch, 11)
        x = torch.from_numpy(g[arch + "_x"]).cuda()
        ref = torch.from_numpy(g[arch + "_y"]).cuda()
        y = m(x)
        p, mx = psnr(y, ref), float((y - ref).abs().max())
        print(f"\n[{arch} golden] PSNR {p:.1f} dB, max-abs {mx:.4g}")
        assert p >= RESTORER_TOL["psnr_db"] and mx <= RESTORER_TOL["max_abs"]


def test_restore_u8_is_truncation_of_forward():
    from oracle import models_oracle as O
    from b200restore import synth
    m, _ = _load("simple_unet", 22)
    img, _ = synth.sign_like_images(2, 64, 64, seed=2)
    img = img.cuda()
    y = m(O.to_tensor_u8(img))
    u8_from_f32 = m.restore_u8(O.to_tensor_u8(img))
    assert torch.equal(u8_from_f32, O.quantize_restored(y))
    # u8 NHWC input (ToTensor fused into the first conv) gives the same bytes as the f32 NCHW entry
    assert torch.equal(m.restore_u8(img), u8_from_f32)


def test_micro_batching_is_invisible():
    from oracle import models_oracle as O
    from b200restore import synth
    m, _ = _load("resunet", 23)
    img, _ = synth.sign_like_images(5, 32, 32, seed=3)
    x = O.to_tensor_u8(img).cuda()
    m.micro_batch = 128
    a = m(x)
    m.micro_batch = 2
    b = m(x)
    assert torch.equal(a, b)


def test_repack_on_load_and_head_swap():
    from b200restore import models, synth
    m, _ = _load("simple_unet", 24)
    x = torch.rand((1, 3, 32, 32), generator=torch.Generator().manual_seed(0)).cuda()
    y1 = m(x)
    buf = io.BytesIO()
    torch.save(synth.synthetic_state_dict("simple_unet", 25), buf)
    buf.seek(0)
    m.load_state_dict(torch.load(buf, map_location="cuda"))          # 17_run_unified_inference.py:63
    y2 = m(x)
    assert not torch.equal(y1, y2)
    j, _ = _load("vgg16", 26)
    xi = torch.randn((1, 3, 64, 64), generator=torch.Generator().manual_seed(1)).cuda()
    l1 = j(xi)
    j.classifier[6] = torch.nn.Linear(4096, 10).cuda()               # the reference's head swap (06:65-67)
    assert j(xi).shape == (1, 10) and l1.shape == (1, 43)


# bf16 activations through 13 convs + 2 FC layers: observed max |logit error| = 2-3 % of the logit standard deviation
JUDGE_TOL = {"rel_to_logit_std": 0.05}


@pytest.mark.parametrize("hw", [(224, 224), (64, 64), (256, 256)])
def test_vgg16_logits_vs_oracle(hw):
    from oracle import models_oracle as O
    from b200restore import synth
    j, sd = _load("vgg16", 27)
    img, _ = synth.sign_like_images(4, hw[0], hw[1], seed=4)
    img = img.cuda()
    x = O.normalize_imagenet(O.to_tensor_u8(img))
    with torch.no_grad():
        ref = O.vgg16_forward(sd, x)
    got = j(x)
    got_u8 = j.forward_u8(img)
    err = float((got - ref).abs().max())
    scale = float(ref.std())
    print(f"\n[vgg16 {hw}] max logit err {err:.4g}, logit std {scale:.4g}, ratio {err / scale:.4g}")
    assert got.shape == (4, 43)
    assert err <= JUDGE_TOL["rel_to_logit_std"] * scale
    # the fused u8 entry (ToTensor + Normalize inside the first conv) == the f32 entry up to bf16 rounding of layer 1
    assert float((got_u8 - got).abs().max()) <= JUDGE_TOL["rel_to_logit_std"] * scale


def test_vgg16_matches_committed_torchvision_output():
    g = golden("models_ref.npz")
    j, _ = _load("vgg16", 13)
    ref = torch.from_numpy(g["vgg16_y"]).cuda()
    got = j(torch.from_numpy(g["vgg16_x"]).cuda())
    err, scale = float((got - ref).abs().max()), float(ref.std())
    print(f"\n[vgg16 golden] max err {err:.4g} vs logit std {scale:.4g}")
    assert err <= JUDGE_TOL["rel_to_logit_std"] * scale


def test_vgg_feature_taps_vs_oracle():
    """Scripts 11 / 12 by-products: channel-mean heat-map of features[:k+1] and the 512-d GAP embedding.
    bf16 activations vs the fp32 oracle.  Tolerances (observed on B200 in parentheses): min-max normalised heat-map
    max-abs <= 0.02 (4e-3); embedding error <= 5 % of its standard deviation over the batch (1-2 %)."""
    from oracle import models_oracle as O
    from b200restore import synth
    m, sd = _load("vgg16", 5)
    img, _ = synth.sign_like_images(3, 96, 128, seed=4)
    img = img.cuda()
    x = O.normalize_imagenet(O.to_tensor_u8(img))
    with torch.no_grad():
        for k in (0, 1, 2, 3, 4, 5, 9):
            ref = O.vgg16_heatmap(sd, x, k)
            got = m.feature_heatmap(img, layer_index=k)
            got_f32 = m.feature_heatmap(x, layer_index=k)
            assert got.shape == ref.shape and got.dtype == torch.float32
            e = float((got - ref).abs().max())
            print(f"\n[heatmap layer {k}] max-abs {e:.4g}", end="")
            assert e <= 0.02 and float((got_f32 - ref).abs().max()) <= 0.02, k
        ref = O.vgg16_gap_embedding(sd, x)
        got = m.feature_embedding(img)
        assert got.shape == (3, 512)
        rel = float((got - ref).abs().max() / ref.std())
        print(f"\n[embedding] max error / std {rel:.4g}")
        assert rel <= 0.05
    with pytest.raises(Exception):
        m.feature_heatmap(img, layer_index=31)


@pytest.mark.parametrize("hw", [(100, 90), (37, 51), (9, 8), (226, 222)])
def test_resunet_sizes_that_need_the_nearest_realignment(hw):
    """H or W not a multiple of 8: the max-pools floor and F.interpolate (nearest) re-aligns every up-sampled tensor to its
    skip connection (14_train_unified_advanced.py:169-183) — round 1 rejected these sizes.  Same bars as the fused path."""
    from b200restore import models, synth
    from oracle import models_oracle as MO
    sd = synth.synthetic_state_dict("resunet", 12)
    m = models.ResUNet()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = torch.rand((2, 3, hw[0], hw[1]), device="cuda")
    with torch.no_grad():
        ref = MO.resunet_forward({k: v.cuda() for k, v in sd.items()}, x)
    out = m(x)
    assert out.shape == ref.shape
    err = float((out - ref).abs().max())
    mse = float(((out - ref) ** 2).mean())
    psnr = 10 * math.log10(float((ref.max() - ref.min()) ** 2) / mse)
    print(f"\n[resunet {hw}] PSNR {psnr:.1f} dB, max-abs {err:.4g}")
    assert psnr >= 50.0 and err <= 2e-2
    u8 = (x.permute(0, 2, 3, 1) * 255).to(torch.uint8).contiguous()
    d = (m.restore_u8(u8).int() - MO.quantize_restored(MO.resunet_forward({k: v.cuda() for k, v in sd.items()}, MO.to_tensor_u8(u8))).int()).abs()
    assert float(d.float().mean()) < 0.5 and int(d.max()) <= 4


@pytest.mark.parametrize("in_c,out_c,hw", [(64, 64, (24, 40)), (128, 256, (28, 28)), (384, 128, (14, 30)), (512, 512, (8, 8))])
def test_residual_block_stand_alone_forward(in_c, out_c, hw):
    """ResidualBlock.forward on its own (14:114-115) against the fp32 oracle; identity and 1x1 + BN shortcuts."""
    from b200restore import models, synth
    from oracle import models_oracle as MO
    full = synth.synthetic_state_dict("resunet", 5)
    name = {(64, 64): "res1", (128, 256): "res3", (384, 128): "dec3", (512, 512): "bottleneck.1"}[(in_c, out_c)]
    sd = {k[len(name) + 1:]: v for k, v in full.items() if k.startswith(name + ".")}
    rb = models.ResidualBlock(in_c, out_c)
    rb.load_state_dict(sd)
    rb = rb.cuda().eval()
    x = torch.randn((3, in_c, hw[0], hw[1]), device="cuda") * 0.5
    xb = x.to(torch.bfloat16).float()                      # what the block sees (NHWC bf16 inside)
    with torch.no_grad():
        ref = MO.residual_block_forward({k: v.cuda() for k, v in full.items()}, name, xb)
    out = rb(x)
    assert out.shape == ref.shape and out.dtype == torch.float32
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    print(f"\n[ResidualBlock {in_c}->{out_c}] max-abs {err:.4g} of range {scale:.3g}")
    assert err <= 2.0 ** -6 * scale + 1e-3
    from b200restore import B2RError
    with pytest.raises(B2RError):
        rb(x[:, :in_c - 1])
