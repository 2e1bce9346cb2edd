"""The driver mirrors of harness.py on the GPU against the reference's own composition of torchvision / PIL calls:
evaluate_model (18_test_unified_benchmark.py:22-53) and run_inference (17_run_unified_inference.py:57-101)."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _tree(root, classes, per_class, exts=("png", "ppm"), seed=0):
    from b200restore import synth
    rng = np.random.default_rng(seed)
    files = []
    k = 0
    for cls in classes:
        (root / cls).mkdir(parents=True)
        for j in range(per_class):
            h, w = int(rng.integers(25, 140)), int(rng.integers(25, 140))     # GTSRB sizes are ragged (15..250 px)
            img, _ = synth.sign_like_images(1, h, w, seed=100 + k)
            f = root / cls / f"{j:05d}.{exts[k % len(exts)]}"
            Image.fromarray(img[0].numpy()).save(f)
            files.append(f)
            k += 1
    return files


def test_evaluate_model_matches_the_reference_composition(tmp_path, capsys):
    from torchvision import datasets, transforms
    from b200restore import harness, imageio as IO, models, synth
    from oracle import models_oracle as MO
    root = tmp_path / "restored" / "Compound"
    _tree(root, ["00000", "00007", "00013"], 5)
    sd = synth.synthetic_state_dict("vgg16", 32)
    judge = models.VGG16Judge()
    judge.load_state_dict(sd)
    judge = judge.cuda().eval()
    acc, pred = harness.evaluate_model(judge, root, "Unified Restored", batch_size=4, return_predictions=True)
    out = capsys.readouterr().out
    assert "Testing: Unified Restored (Total 15 images)..." in out and f"-> Unified Restored Accuracy: {acc * 100:.2f}%" in out
    # oracle: exactly the reference's loader (18:28-36) feeding the fp32 judge
    tf = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ds = datasets.ImageFolder(root=str(root), transform=tf)
    x = torch.stack([ds[i][0] for i in range(len(ds))]).cuda()
    labels = torch.tensor([ds[i][1] for i in range(len(ds))]).cuda()
    with torch.no_grad():
        logits_ref = MO.vgg16_forward({k: v.cuda() for k, v in sd.items()}, x)
    pred_ref = logits_ref.argmax(1)
    logits = judge.forward_u8(IO.load_batch([p for p, _ in ds.samples]))
    lerr = float((logits - logits_ref).abs().max())
    top2 = torch.topk(logits_ref, 2, dim=1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > 2 * lerr
    assert bool((pred == pred_ref)[safe].all())
    assert lerr <= 0.05 * float(logits_ref.std()) + 1e-3
    assert acc == pytest.approx(float((pred == labels).float().mean()), abs=1e-12)   # the count is that of OUR labels
    assert harness.benchmark_table(judge, {"Unified Restored": root, "Missing": tmp_path / "nope"}, verbose=False) == \
        {"Unified Restored": acc}


def test_run_inference_writes_the_restored_tree(tmp_path):
    from b200restore import harness, models, synth
    from oracle import imageio_oracle as IOO, models_oracle as MO
    src, dst = tmp_path / "processed" / "Compound", tmp_path / "restored" / "Compound"
    files = _tree(src, ["00001", "00042"], 5, exts=("png",), seed=1)
    sd = synth.synthetic_state_dict("resunet", 31)
    m = models.ResUNet()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    written = harness.run_inference(m, src, dst, batch_size=4, verbose=False)
    assert sorted(p.relative_to(dst) for p in written) == sorted(f.relative_to(src) for f in files)
    order = sorted(files)
    with torch.no_grad():
        ref = MO.quantize_restored(MO.resunet_forward({k: v.cuda() for k, v in sd.items()},
                                                      MO.to_tensor_u8(torch.from_numpy(IOO.load_and_resize(order)).cuda())))
    got = np.stack([np.asarray(Image.open(dst / f.relative_to(src))) for f in order])
    assert got.shape == (len(files), 224, 224, 3)
    d = np.abs(got.astype(int) - ref.cpu().numpy().astype(int))
    assert d.mean() < 0.5 and (d > 2).mean() < 1e-3, (float(d.mean()), int(d.max()))
    with pytest.raises(Exception):
        harness.run_inference(models.ResUNet(), src, dst, verbose=False)           # module on the CPU: no fallback


def test_process_task_scores_like_script_08(tmp_path, capsys):
    """08:55-137: distorted tree -> restored .png tree + mean PSNR / SSIM against cv2.resize(clean, (224, 224)).
    Oracle: the reference's own composition on the files this run wrote (cv2.resize + the oracle's PSNR / SSIM)."""
    from b200restore import harness, models, synth
    from oracle import generators_oracle as GO, imageio_oracle as IOO
    dist, rest, clean = tmp_path / "processed" / "Noise", tmp_path / "restored" / "Noise", tmp_path / "gtsrb"
    files = _tree(dist, ["00003", "00021"], 4, exts=("png", "ppm"), seed=2)
    rng = np.random.default_rng(5)
    for k, f in enumerate(files):                       # clean counterparts: other sizes, .ppm names; one missing
        if k == 3:
            continue
        cp = (clean / f.relative_to(dist)).with_suffix(".ppm")
        cp.parent.mkdir(parents=True, exist_ok=True)
        h, w = int(rng.integers(20, 200)), int(rng.integers(20, 200))
        Image.fromarray(synth.sign_like_images(1, h, w, seed=300 + k)[0][0].numpy()).save(cp)
    m = models.SimpleUNet()
    m.load_state_dict(synth.synthetic_state_dict("simple_unet", 41))
    m = m.cuda().eval()
    res = harness.process_task(m, "Noise", dist, rest, clean, batch_size=3)
    out = capsys.readouterr().out
    assert res is not None and res[2] == len(files) - 1
    assert "=== Starting task processing: Noise ===" in out and f"Average PSNR: {res[0]:.2f} dB" in out
    ps, ss = [], []
    for k, f in enumerate(files):
        written = (rest / f.relative_to(dist)).with_suffix(".png")
        assert written.exists()
        if k == 3:
            continue
        ref = IOO.resize_cv(np.asarray(Image.open((clean / f.relative_to(dist)).with_suffix(".ppm")).convert("RGB")))
        got = np.asarray(Image.open(written))
        ps.append(GO.psnr_08(ref, got))
        ss.append(GO.ssim_08(ref, got))
    assert res[0] == pytest.approx(float(np.mean(ps)), rel=1e-12)
    assert res[1] == pytest.approx(float(np.mean(ss)), abs=1e-10)
    assert harness.process_task(m, "Fog", tmp_path / "processed" / "Fog", rest, clean, verbose=False) is None
