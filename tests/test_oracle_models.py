"""Pin the model oracle and the checkpoint schema: functional fp32 restatement vs the reference's nn.Module classes
(live, when /root/reference is mounted) and vs the committed reference outputs; synthetic state_dicts load strictly
into the reference classes and round-trip through torch.save / torch.load into the drop-in modules."""
import io

import numpy as np
import pytest
import torch

from _util import golden, have_reference, load_ref
from oracle import models_oracle as O


def _sd(arch, seed):
    from b200restore import synth
    return synth.synthetic_state_dict(arch, seed)


@pytest.mark.parametrize("arch,fn,seed", [("simple_unet", O.simple_unet_forward, 11), ("resunet", O.resunet_forward, 11),
                                          ("vgg16", O.vgg16_forward, 13)])
def test_oracle_matches_golden_reference_outputs(arch, fn, seed):
    g = golden("models_ref.npz")
    sd = _sd(arch, seed)
    # the seeded synthetic checkpoint is the one the fixture was generated with
    cs = dict(zip(g[arch + "_cs_keys"].tolist(), g[arch + "_cs_vals"].tolist()))
    for k, v in sd.items():
        assert abs(float(v.double().abs().sum()) - cs[k]) <= 1e-9 * max(1.0, abs(cs[k])), k
    with torch.no_grad():
        y = fn(sd, torch.from_numpy(g[arch + "_x"]))
    ref = torch.from_numpy(g[arch + "_y"])
    # same ATen kernels, same order of operations: equal up to thread-count dependent summation order
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5), float((y - ref).abs().max())


@pytest.mark.skipif(not have_reference(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference_modules():
    r07, r14, r17 = (load_ref(f) for f in ("07_train_restoration.py", "14_train_unified_advanced.py",
                                           "17_run_unified_inference.py"))
    g = torch.Generator().manual_seed(0)
    x = torch.rand((1, 3, 32, 40), generator=g)
    with torch.no_grad():
        for ctor, arch, fn in ((r07.SimpleUNet, "simple_unet", O.simple_unet_forward),
                               (r14.ResUNet, "resunet", O.resunet_forward), (r17.ResUNet, "resunet", O.resunet_forward)):
            sd = _sd(arch, 3)
            m = ctor()
            m.load_state_dict(sd, strict=True)
            m.eval()
            assert torch.equal(m(x), fn(sd, x)), arch
        # ResUNet's nearest-neighbour re-alignment branch (14:169-183): H, W not multiples of 8
        sd = _sd("resunet", 4)
        m = r14.ResUNet()
        m.load_state_dict(sd)
        m.eval()
        xo = torch.rand((1, 3, 36, 44), generator=g)
        assert torch.equal(m(xo), O.resunet_forward(sd, xo))


def test_vgg16_oracle_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    sd = _sd("vgg16", 5)
    m = tv.models.vgg16(weights=None)
    m.classifier[6] = torch.nn.Linear(m.classifier[6].in_features, 43)     # 06_test_baseline.py:65-67
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = torch.randn((1, 3, 96, 64), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert torch.equal(m(x), O.vgg16_forward(sd, x))


@pytest.mark.skipif(not have_reference(), reason="/root/reference not mounted")
def test_schema_identical_to_reference_classes():
    r07, r14 = load_ref("07_train_restoration.py"), load_ref("14_train_unified_advanced.py")
    from b200restore import models
    for ref_ctor, ours in ((r07.SimpleUNet, models.SimpleUNet), (r14.ResUNet, models.ResUNet)):
        a, b = ref_ctor().state_dict(), ours().state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k


def test_schema_facts_from_survey():
    """26 tensors in SimpleUNet, 1 862 979 / 12 625 869 / 134 436 715 parameters."""
    from b200restore import models
    u, r, v = models.SimpleUNet(), models.ResUNet(), models.VGG16Judge()
    assert len(u.state_dict()) == 26
    assert sum(p.numel() for p in u.parameters()) == 1_862_979
    assert sum(p.numel() for p in r.parameters()) == 12_625_869
    assert sum(b.numel() for b in r.buffers()) == 10_777
    assert sum(p.numel() for p in v.parameters()) == 134_436_715
    assert tuple(v.state_dict()["classifier.6.weight"].shape) == (43, 4096)
    assert "res1.shortcut.0.weight" not in r.state_dict() and "res2.shortcut.0.weight" in r.state_dict()
    assert r.state_dict()["res2.conv_block.1.num_batches_tracked"].dtype == torch.int64


@pytest.mark.parametrize("arch", ["simple_unet", "resunet", "vgg16"])
def test_checkpoint_roundtrip_strict_load(arch):
    """torch.save(model.state_dict(), path) -> load_state_dict(torch.load(path, map_location=...)) strict, as
    17_run_unified_inference.py:63 does, including int64 num_batches_tracked; a wrong key raises RuntimeError."""
    from b200restore import models
    sd = _sd(arch, 9)
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    loaded = torch.load(buf, map_location="cpu")
    m = {"simple_unet": models.SimpleUNet, "resunet": models.ResUNet, "vgg16": models.VGG16Judge}[arch]()
    res = m.load_state_dict(loaded)            # strict=True is the default
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    bad = dict(loaded)
    bad["not.a.key"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)


def test_modules_refuse_cpu_and_training_mode():
    from b200restore import models, B2RError
    m = models.SimpleUNet()
    with pytest.raises(B2RError):
        m(torch.zeros(1, 3, 16, 16))           # training mode
    m.eval()
    with pytest.raises(B2RError):
        m(torch.zeros(1, 3, 16, 16))           # CPU tensor: no fallback


def test_pipeline_composition_shapes():
    sdr, sdj = _sd("simple_unet", 1), _sd("vgg16", 2)
    img = torch.randint(0, 256, (2, 32, 32, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    restored, logits, pred = O.restore_then_classify(O.simple_unet_forward, sdr, sdj, img)
    assert restored.dtype == torch.uint8 and restored.shape == (2, 32, 32, 3)
    assert logits.shape == (2, 43) and pred.dtype == torch.int64
    # truncation, not rounding (17_run_unified_inference.py:92)
    assert torch.equal(O.quantize_restored(torch.full((1, 3, 1, 1), 0.999)), torch.full((1, 1, 1, 3), 254, dtype=torch.uint8))


def test_vgg_feature_taps_match_torchvision_slicing():
    """oracle taps == torchvision's own `model.features[:k + 1]` on the same seeded weights (what scripts 11 / 12 run)."""
    import torchvision
    from b200restore import synth
    from oracle import models_oracle as O
    sd = synth.synthetic_state_dict("vgg16", 3)
    tv = torchvision.models.vgg16(weights=None)
    tv.classifier[6] = torch.nn.Linear(4096, 43)
    tv.load_state_dict(sd)
    tv.eval()
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for k in (0, 1, 2, 3, 4, 5, 16, 30):
            assert torch.equal(O.vgg16_features_upto(sd, x, k), tv.features[:k + 1](x)), k
        assert torch.equal(O.vgg16_gap_embedding(sd, x), torch.mean(tv.features(x), dim=[2, 3]))
        h = torch.mean(tv.features[:3](x), dim=1)
        ref = torch.stack([(h[i] - h[i].min()) / (h[i].max() - h[i].min()) for i in range(2)])
        assert torch.allclose(O.vgg16_heatmap(sd, x, 2), ref, atol=0, rtol=0)
