"""Argument guards added after the round-1 review: device mismatch, predictions-only pipeline calls (eager and graph),
in-place blur, explicit re-pack after a `.data` write."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pipe(mb=8):
    from b200restore import models, synth, RestoreClassifyPipeline
    r, j = models.SimpleUNet(), models.VGG16Judge()
    r.load_state_dict(synth.synthetic_state_dict("simple_unet", 31))
    j.load_state_dict(synth.synthetic_state_dict("vgg16", 32))
    return RestoreClassifyPipeline(r.cuda(), j.cuda(), micro_batch=mb)


def test_predictions_only_runs_in_eager_and_graph_mode():
    from b200restore import degrade, synth
    pipe = _pipe()
    n = 12
    imgs, labels = synth.indexed_images(0, n, 64, 64, seed=9)
    imgs, labels = imgs.cuda(), labels.cuda()
    params = degrade.compound_params(n)
    p_ref, c_ref = pipe.run(imgs, labels, params, seed=5)
    for graph in (False, True):
        pipe.use_graph = graph
        p, c = pipe.run(imgs, None, params, seed=5)             # labels = None: predictions + total only
        assert torch.equal(p, p_ref) and c.tolist() == [0, n]
        p2, c2 = pipe.run(imgs, labels, params, seed=5)
        assert torch.equal(p2, p_ref) and torch.equal(c2, c_ref)


def test_degrade_rejects_in_place_blur_and_allows_in_place_pointwise():
    from b200restore import _lib as L, degrade, synth
    imgs, _ = synth.indexed_images(0, 4, 64, 64, seed=1)
    x = imgs.cuda()
    with pytest.raises(L.B2RError, match="overlap"):
        degrade.degrade(x, degrade.compound_params(4), out=x)
    ref = degrade.degrade(x, degrade.fog_params(4, __import__("numpy").random.default_rng(0)))
    y = x.clone()
    degrade.degrade(y, degrade.fog_params(4, __import__("numpy").random.default_rng(0)), out=y)   # no blur: in place is fine
    assert torch.equal(y, ref)


def test_invalidate_pack_after_data_write():
    from b200restore import models, synth
    m = models.SimpleUNet()
    m.load_state_dict(synth.synthetic_state_dict("simple_unet", 3))
    m = m.cuda().eval()
    x = torch.rand(2, 3, 32, 32, device="cuda")
    y0 = m(x)
    m.final.bias.data.add_(0.25)                                # invisible to the version counter
    m.invalidate_pack()
    y1 = m(x)
    assert torch.allclose(y1, y0 + 0.25, atol=1e-5)
    with torch.no_grad():
        m.final.bias.sub_(0.25)                                  # visible: re-packed automatically
    assert torch.allclose(m(x), y0, atol=1e-5)


def test_wrong_current_device_is_rejected():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from b200restore import _lib as L, ops
    x = torch.zeros((1, 8, 8, 64), dtype=torch.bfloat16, device="cuda:1")
    with torch.cuda.device(0), pytest.raises(L.B2RError, match="current CUDA device"):
        ops.maxpool2x2(x)
    with torch.cuda.device(1):
        ops.maxpool2x2(x)


def test_modules_enter_their_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from b200restore import models, synth
    m = models.SimpleUNet()
    m.load_state_dict(synth.synthetic_state_dict("simple_unet", 3))
    x = torch.rand(2, 3, 32, 32)
    y0 = m.cuda(0).eval()(x.cuda(0))
    with torch.cuda.device(0):
        y1 = m.cuda(1)(x.cuda(1))                                # current device 0, module and input on 1
    assert torch.equal(y0.cpu(), y1.cpu())
