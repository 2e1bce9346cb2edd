"""GPU parity of the on-disk edge (imageio.py + b2r_resize_bilinear_u8): ragged batches resized on the device must equal
Pillow / torchvision's Resize((224, 224)) byte for byte, files must round-trip through the reference's tree layout."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


def test_resize_batch_bit_exact_on_ragged_batches():
    from b200restore import imageio as IO
    from oracle import imageio_oracle as IOO
    rng = np.random.default_rng(2)
    sizes = [(224, 224), (15, 15), (250, 250), (224, 61), (97, 224), (30, 29), (1000, 40), (40, 900), (225, 223)]
    sizes += [(int(rng.integers(15, 260)), int(rng.integers(15, 260))) for _ in range(40)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    got = IO.resize_batch(imgs).cpu().numpy()
    assert got.shape == (len(imgs), 224, 224, 3)
    for i, im in enumerate(imgs):
        assert np.array_equal(got[i], IOO.resize_pil(im)), sizes[i]
    # another output size, smooth content, single image
    smooth = np.clip(np.add.outer(np.arange(120), np.arange(77))[:, :, None] + np.array([0, 40, 90]), 0, 255).astype(np.uint8)
    assert np.array_equal(IO.resize_batch([smooth], size=(64, 96)).cpu().numpy()[0], IOO.resize_pil(smooth, (64, 96)))
    assert IO.resize_batch([]).shape == (0, 224, 224, 3)
    with pytest.raises(Exception):
        IO.resize_batch([np.zeros((4, 4), np.uint8)])


def test_load_restore_save_tree_roundtrip(tmp_path):
    """17:69-99 on a small tree: class folders with .ppm / .png files -> device batch -> files under another root."""
    from b200restore import imageio as IO
    from oracle import imageio_oracle as IOO
    rng = np.random.default_rng(3)
    src, dst = tmp_path / "distorted", tmp_path / "restored"
    files = []
    for cls in ("00000", "00017"):
        (src / cls).mkdir(parents=True)
        for k in range(3):
            h, w = int(rng.integers(20, 120)), int(rng.integers(20, 120))
            f = src / cls / f"{k:05d}.{'ppm' if k % 2 else 'png'}"
            Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(f)
            files.append(f)
    batch = IO.load_batch(files)
    assert np.array_equal(batch.cpu().numpy(), IOO.load_and_resize(files))
    out = IO.save_batch(batch, files, src, dst, suffix=".png")
    assert [p.relative_to(dst).with_suffix("") for p in out] == [f.relative_to(src).with_suffix("") for f in files]
    for p, ref in zip(out, batch.cpu().numpy()):
        assert np.array_equal(np.asarray(Image.open(p)), ref)


def test_resize_batch_cv_bit_exact_against_cv2():
    """cv2.resize(img, (224, 224)) (08:119) on a ragged batch in one launch: bit-exact against OpenCV itself, including
    up-scaling from 15 px, down-scaling from 600 px, the identity size and a non-square target."""
    from b200restore import imageio as IO
    from oracle import imageio_oracle as IOO
    rng = np.random.default_rng(4)
    sizes = [(15, 15), (15, 250), (250, 15), (224, 224), (448, 448), (600, 333), (31, 97), (100, 100), (223, 225), (1, 7)]
    sizes += [(int(rng.integers(15, 251)), int(rng.integers(15, 251))) for _ in range(30)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    out = IO.resize_batch_cv(imgs).cpu().numpy()
    for i, im in enumerate(imgs):
        assert np.array_equal(out[i], IOO.resize_cv(im)), sizes[i]
    out2 = IO.resize_batch_cv(imgs[:12], size=(96, 160)).cpu().numpy()
    for i in range(12):
        assert np.array_equal(out2[i], IOO.resize_cv(imgs[i], (96, 160))), sizes[i]
    assert IO.resize_batch_cv([]).shape == (0, 224, 224, 3)
