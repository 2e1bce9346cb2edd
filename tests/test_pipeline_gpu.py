"""End-to-end degrade -> restore -> quantise -> VGG16 -> top-1 -> count on the GPU against the oracle pipeline
(15_test_unified.py:170-200 semantics) on identical images, parameters and injected noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _build(restorer_arch):
    from b200restore import models, synth, RestoreClassifyPipeline
    sdr = synth.synthetic_state_dict(restorer_arch, 31)
    sdj = synth.synthetic_state_dict("vgg16", 32)
    r = (models.SimpleUNet if restorer_arch == "simple_unet" else models.ResUNet)()
    r.load_state_dict(sdr)
    j = models.VGG16Judge()
    j.load_state_dict(sdj)
    pipe = RestoreClassifyPipeline(r.cuda(), j.cuda(), micro_batch=8)
    return pipe, {k: v.cuda() for k, v in sdr.items()}, {k: v.cuda() for k, v in sdj.items()}


@pytest.mark.parametrize("arch", ["simple_unet", "resunet"])
def test_pipeline_vs_oracle(arch):
    from b200restore import degrade, synth
    from oracle import degrade_oracle as DO, models_oracle as MO
    pipe, sdr, sdj = _build(arch)
    n, h, w = 12, 224, 224
    imgs, labels = synth.sign_like_images(n, h, w, seed=7)
    z = np.random.default_rng(8).standard_normal((n, h, w, 3))
    noise = (0.02 ** 0.5) * z
    # oracle: script 16 degradation per image, then the 15:179-200 composition in fp32 on the GPU
    deg_ref = np.stack([DO.compound_16(imgs[i].numpy(), noise[i]) for i in range(n)])
    fn = MO.simple_unet_forward if arch == "simple_unet" else MO.resunet_forward
    restored_ref, logits_ref, pred_ref = MO.restore_then_classify(fn, sdr, sdj, torch.from_numpy(deg_ref).cuda())
    # product: one micro-batch with everything kept for inspection
    params = degrade.compound_params(n)
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    pipe.micro_batch = n
    pred, extra = pipe.run_micro_batch(imgs.cuda(), labels.cuda(), params.to("cuda"), 0, 0, counts,
                                       noise=torch.from_numpy(noise).cuda(), keep=True)
    assert np.array_equal(extra["degraded"].cpu().numpy(), deg_ref)                 # byte work: bit-exact
    d = (extra["restored"].int() - restored_ref.int()).abs()
    print(f"\n[{arch}] restored u8: max |diff| {int(d.max())}, mean {float(d.float().mean()):.4f}, "
          f"frac>1 {float((d > 1).float().mean()):.2e}")
    assert float(d.float().mean()) < 0.5 and float((d > 2).float().mean()) < 1e-3   # bf16 restorer, u8 LSBs
    margin = torch.topk(logits_ref, 2, dim=1)[0]
    margin = (margin[:, 0] - margin[:, 1])
    agree = (pred == pred_ref)
    lerr = float((extra["logits"] - logits_ref).abs().max())
    print(f"[{arch}] top-1 agreement {int(agree.sum())}/{n}; max logit err {lerr:.4g}; "
          f"ref margins min {float(margin.min()):.4g} median {float(margin.median()):.4g}")
    # labels must agree wherever the reference's own top-1/top-2 margin exceeds twice the observed logit error
    safe = margin > 2 * lerr
    assert bool(agree[safe].all())
    assert counts[1].item() == n and counts[0].item() == int((pred == labels.cuda()).sum())


def test_run_and_run_from_host_agree_and_count():
    from b200restore import degrade, synth
    pipe, _, _ = _build("simple_unet")
    n = 20
    imgs, labels = synth.sign_like_images(n, 64, 64, seed=9)
    params = degrade.compound_params(n)
    pred, counts = pipe.run(imgs.cuda(), labels.cuda(), params, seed=5, image_index0=100)
    (correct, total), h2d, d2h = pipe.run_from_host(imgs.pin_memory(), labels.pin_memory(), params, seed=5,
                                                    image_index0=100)
    assert total == n == counts[1].item() and correct == counts[0].item()
    assert correct == int((pred == labels.cuda()).sum())
    assert h2d == n * 64 * 64 * 3 + n * 8 and d2h == 16
    # micro-batch size and launch partition do not change any prediction (noise is keyed by the global image index)
    pipe.micro_batch = 3
    pred2, counts2 = pipe.run(imgs.cuda(), labels.cuda(), params, seed=5, image_index0=100)
    assert torch.equal(pred, pred2) and torch.equal(counts, counts2)
