"""GPU parity of the CUDA-core kernels around the convs, each against PyTorch fp32 (or exact integer semantics)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("n,h,w", [(2, 16, 16), (1, 37, 53), (3, 224, 224)])
def test_conv3x3_c3_f32_nchw(n, h, w):
    from b200restore import ops, _lib as L
    x = torch.rand((n, 3, h, w), generator=_g(0)).cuda()
    wt = (torch.randn((64, 3, 3, 3), generator=_g(1)) * 0.3).cuda()
    b = (torch.randn(64, generator=_g(2)) * 0.1).cuda()
    from b200restore import packing
    out = ops.conv3x3_c3(x, packing.pack_conv_c3(wt), b, act=L.B2R_ACT_PRELU, slope=0.2)
    # the kernel feeds each fp32 input as a (hi, lo) bf16 pair (~16 mantissa bits) against bf16 weights, fp32 accumulate:
    # vs torch fp32 on the SAME bf16-rounded weights only the output's bf16 rounding (2^-9) and ~1e-4 of input split remain
    ref = F.prelu(F.conv2d(x, wt.to(torch.bfloat16).float(), b, padding=1), torch.tensor([0.2], device="cuda"))
    err = (nchw(out) - ref).abs()
    assert bool((err <= 2.0 ** -8 * ref.abs() + 2e-4).all()), float(err.max())


def test_conv3x3_c3_u8_normalized():
    from b200restore import ops, _lib as L
    from oracle import models_oracle as O
    u8 = torch.randint(0, 256, (2, 40, 48, 3), dtype=torch.uint8, generator=_g(3)).cuda()
    wt = (torch.randn((64, 3, 3, 3), generator=_g(4)) * 0.3).cuda()
    b = (torch.randn(64, generator=_g(5)) * 0.1).cuda()
    from b200restore import packing
    wp, wr = packing.pack_conv_c3(wt), wt.to(torch.bfloat16).float()
    out = ops.conv3x3_c3(u8, wp, b, act=L.B2R_ACT_RELU, normalize=True)
    ref = F.relu(F.conv2d(O.normalize_imagenet(O.to_tensor_u8(u8)), wr, b, padding=1))
    err = (nchw(out) - ref).abs()
    assert bool((err <= 2.0 ** -8 * ref.abs() + 5e-4).all()), float(err.max())
    out2 = ops.conv3x3_c3(u8, wp, b, act=L.B2R_ACT_RELU, normalize=False)
    ref2 = F.relu(F.conv2d(O.to_tensor_u8(u8), wr, b, padding=1))
    err2 = (nchw(out2) - ref2).abs()
    assert bool((err2 <= 2.0 ** -8 * ref2.abs() + 2e-4).all()), float(err2.max())


def test_final_conv1x1_f32_and_quantised_u8():
    from b200restore import ops
    from oracle import models_oracle as O
    x = (torch.randn((2, 24, 40, 64), generator=_g(6)) * 0.5).to(torch.bfloat16).cuda()
    wt = (torch.randn((3, 64, 1, 1), generator=_g(7)) * 0.2).cuda()
    b = torch.tensor([0.4, 0.5, 0.6]).cuda()
    o32, o8 = ops.final_conv1x1(x, wt, b, want_f32=True, want_u8=True)
    ref = F.conv2d(nchw(x), wt, b)
    assert torch.allclose(o32, ref, rtol=1e-5, atol=1e-5)
    # the u8 output is the reference's clamp -> *255 -> truncation applied to THIS kernel's own f32 output
    assert torch.equal(o8, O.quantize_restored(o32))
    # and differs from the quantised torch result only where the two f32 results straddle an integer boundary
    d = (o8.int() - O.quantize_restored(ref).int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3


def test_maxpool_and_avgpool():
    from b200restore import ops
    x = torch.randn((3, 12, 20, 128), generator=_g(8)).to(torch.bfloat16).cuda()
    assert torch.equal(nchw(ops.maxpool2x2(x)), F.max_pool2d(nchw(x), 2, 2))
    for h, w in ((7, 7), (2, 2), (8, 8), (14, 9)):
        y = torch.randn((2, h, w, 64), generator=_g(9)).to(torch.bfloat16).cuda()
        got = nchw(ops.adaptive_avgpool7(y))
        ref = F.adaptive_avg_pool2d(nchw(y), (7, 7))
        assert torch.allclose(got, ref, rtol=2.0 ** -8, atol=1e-6), (h, w)


def test_linear_f32out():
    from b200restore import ops
    x = torch.randn((37, 4096), generator=_g(10)).to(torch.bfloat16).cuda()
    w = (torch.randn((43, 4096), generator=_g(11)) * 0.02).to(torch.bfloat16).cuda()
    b = torch.randn(43, generator=_g(12)).cuda()
    got = ops.linear_f32out(x, w, b)
    ref = x.float() @ w.float().t() + b
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)


def test_argmax_count_matches_torch_max_including_ties():
    from b200restore import ops
    logits = torch.randn((1001, 43), generator=_g(13)).cuda()
    logits[5, :] = 1.0                      # all equal: torch.max returns index 0
    logits[6, 40] = logits[6, 3] = 99.0     # tie: lowest index wins
    logits[7, 42] = 50.0                    # last column
    labels = torch.randint(0, 43, (1001,), generator=_g(14)).cuda()
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    pred, conf = ops.argmax_count(logits, labels, counts, want_conf=True)
    ref_v, ref_i = torch.max(logits, 1)
    assert torch.equal(pred, ref_i)
    assert pred[5].item() == 0 and pred[6].item() == 3 and pred[7].item() == 42
    assert counts.tolist() == [int((ref_i == labels).sum()), 1001]
    ref_conf = torch.softmax(logits, 1).max(1)[0]            # 15_test_unified.py:125-129
    assert torch.allclose(conf, ref_conf, rtol=1e-5, atol=1e-6)
    pred2, _ = ops.argmax_count(logits, labels, counts)      # counters accumulate across calls
    assert counts.tolist() == [2 * int((ref_i == labels).sum()), 2002]
