"""Whole-network C entry points (include/b2r.h section 7, csrc/net_plan.cu) against the nn.Module path: both issue the same
kernels on the same bytes, so outputs must be BIT-IDENTICAL — this pins the C++ restatement of the weight packing (fp64 BN
fold, k-block order, tap-folded copies, ConvTranspose / classifier layouts) to packing.py, and the C layer graphs to
models.py.  Error behaviour mirrors strict load_state_dict (17:63)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch,hw", [("simple_unet", (64, 96)), ("resunet", (64, 96)), ("resunet", (224, 224))])
def test_restorer_plan_equals_module(arch, hw):
    from b200restore import NetPlan, models, synth
    sd = synth.stress_state_dict(arch, 5) if hw[0] == 64 else synth.synthetic_state_dict(arch, 31)
    m = (models.SimpleUNet if arch == "simple_unet" else models.ResUNet)()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    plan = NetPlan(arch, sd, "cuda")
    n = 3
    imgs, _ = synth.indexed_images(0, n, hw[0], hw[1], seed=4)
    u8 = imgs.cuda()
    x = torch.rand((n, 3, hw[0], hw[1]), device="cuda")
    o32, o8 = plan.restore(u8, want_f32=True, want_u8=True)
    assert torch.equal(o8, m.restore_u8(u8)) and torch.equal(o32, m(u8))
    o32b, _ = plan.restore(x)
    assert torch.equal(o32b, m(x))
    plan.close()


@pytest.mark.parametrize("hw,n", [((224, 224), 5), ((64, 64), 5), ((224, 224), 70), ((64, 64), 800)])
def test_vgg_plan_equals_module(hw, n):
    """n = 70 at 224 x 224 and n = 800 at 64 x 64 take the path that alternates conv1_1 / conv1_2 over sub-batches of 32 x 224^2
    pixels (32 / 392 images; both hosts), which must also equal the whole-batch launches bit for bit (first_stage_sub = 0)."""
    from b200restore import NetPlan, models, synth
    sd = synth.synthetic_state_dict("vgg16", 32)
    j = models.VGG16Judge()
    j.load_state_dict(sd)
    j = j.cuda().eval()
    plan = NetPlan("vgg16", sd, "cuda")
    imgs, _ = synth.indexed_images(10, n, hw[0], hw[1], seed=4)
    u8 = imgs.cuda()
    ref = j.forward_u8(u8)
    assert torch.equal(plan.classify(u8), ref)
    j.first_stage_sub = 0
    assert torch.equal(j.forward_u8(u8), ref)
    j.first_stage_sub = 32
    x = torch.randn((2, 3, hw[0], hw[1]), device="cuda")
    assert torch.equal(plan.classify(x), j(x))
    plan.close()


def test_plan_errors_are_strict_like_load_state_dict():
    from b200restore import NetPlan, _lib as L, synth
    sd = synth.synthetic_state_dict("resunet", 1)
    bad = dict(sd)
    del bad["dec2.shortcut.1.running_var"]
    with pytest.raises(L.B2RError, match="Missing key.*dec2.shortcut.1.running_var"):
        NetPlan("resunet", bad, "cuda")
    bad = dict(sd)
    bad["up1.weight"] = bad["up1.weight"][:, :32].contiguous()
    with pytest.raises(L.B2RError, match="size mismatch for up1.weight"):
        NetPlan("resunet", bad, "cuda")
    plan = NetPlan("resunet", sd, "cuda")
    x = torch.zeros((1, 60, 64, 3), dtype=torch.uint8, device="cuda")          # H not a multiple of 8
    with pytest.raises(L.B2RError, match="multiples of 8"):
        plan.restore(x)
    lib = L.load()
    out = torch.empty((1, 64, 64, 3), dtype=torch.uint8, device="cuda")
    x = torch.zeros((1, 64, 64, 3), dtype=torch.uint8, device="cuda")
    ws = torch.empty(4096, dtype=torch.uint8, device="cuda")                    # far too small
    rc = lib.b2r_resunet_forward(plan.handle, x.data_ptr(), L.B2R_IN_U8_NHWC, None, out.data_ptr(), 1, 64, 64, ws.data_ptr(), 4096, None)
    assert rc == -22 and b"workspace" in lib.b2r_last_error()
    rc = lib.b2r_unet_forward(plan.handle, x.data_ptr(), L.B2R_IN_U8_NHWC, None, out.data_ptr(), 1, 64, 64, ws.data_ptr(), 4096, None)
    assert rc == -22 and b"architecture" in lib.b2r_last_error()
    plan.close()
