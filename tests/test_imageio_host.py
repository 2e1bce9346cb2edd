"""Host side of the on-disk edge: the restated Pillow coefficient tables and the PPM parser against Pillow / torchvision
themselves (the reference's calls, oracle/imageio_oracle.py)."""
import numpy as np
import pytest
from PIL import Image

from oracle import imageio_oracle as IOO


def _apply_tables(img, oh, ow):
    """The arithmetic of csrc/generators.cu::resize_bilinear_u8_kernel in NumPy, on the product's tables."""
    from b200restore import imageio as IO
    h, w, _ = img.shape
    bx, kx = IO.resample_table(w, ow)
    by, ky = IO.resample_table(h, oh)
    tmp = np.zeros((h, ow, 3), np.uint8)
    for xx in range(ow):
        lo, cnt = bx[xx]
        ss = (1 << 21) + (img[:, lo:lo + cnt].astype(np.int64) * kx[xx, :cnt][None, :, None]).sum(1)
        tmp[:, xx] = np.clip(ss >> 22, 0, 255)
    out = np.zeros((oh, ow, 3), np.uint8)
    for yy in range(oh):
        lo, cnt = by[yy]
        ss = (1 << 21) + (tmp[lo:lo + cnt].astype(np.int64) * ky[yy, :cnt][:, None, None]).sum(0)
        out[yy] = np.clip(ss >> 22, 0, 255)
    return out


def test_tables_reproduce_pillow_and_torchvision_resize():
    rng = np.random.default_rng(0)
    sizes = [(224, 224), (224, 100), (15, 15), (250, 250), (31, 222), (640, 37), (225, 223)]
    sizes += [(int(rng.integers(15, 260)), int(rng.integers(15, 260))) for _ in range(12)]
    for h, w in sizes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = IOO.resize_pil(img)
        assert np.array_equal(ref, np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR)))
        assert np.array_equal(_apply_tables(img, 224, 224), ref), (h, w)
    img = rng.integers(0, 256, (50, 70, 3), dtype=np.uint8)
    assert np.array_equal(_apply_tables(img, 64, 96), IOO.resize_pil(img, (64, 96)))


def test_ppm_reader_matches_pillow(tmp_path):
    from b200restore import imageio as IO
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (33, 45, 3), dtype=np.uint8)
    p = tmp_path / "a.ppm"
    Image.fromarray(img).save(p)
    assert np.array_equal(IO.read_ppm(p), img) and np.array_equal(IO.load_rgb(p), img)
    q = tmp_path / "b.ppm"                                   # header with a comment and odd whitespace
    q.write_bytes(b"P6\n# made by hand\n45  33\n255\n" + img.tobytes())
    assert np.array_equal(IO.read_ppm(q), img)
    g = tmp_path / "gray.png"
    Image.fromarray(img[:, :, 0]).save(g)                    # .convert('RGB') replicates the single channel
    assert np.array_equal(IO.load_rgb(g), np.repeat(img[:, :, :1], 3, 2))
    with pytest.raises(Exception):
        IO.read_ppm(g)


def test_cv_linear_tables_reproduce_cv2_resize():
    """08:119 cv2.resize(clean, (224, 224)): the host tables (imageio.cv_linear_table + the border rule of
    resize_batch_cv) driven through a NumPy model of csrc/generators.cu::resize_cv_linear_u8_kernel give cv2's bytes for
    up- and down-scaling, 1-pixel extents and the 2x case OpenCV reroutes to INTER_AREA."""
    import cv2
    from b200restore import imageio as IO

    def model(img, oh, ow):
        h, w, _ = img.shape
        xt, yt = IO.cv_linear_table(w, ow).copy(), IO.cv_linear_table(h, oh)
        outside = (xt[:, 0] < 0) | (xt[:, 0] >= w - 1)
        xt[outside, 1], xt[outside, 2] = 2048, 0
        xt[:, 0] = np.clip(xt[:, 0], 0, w - 1)
        S = img.astype(np.int64)
        sx = xt[:, 0]
        x1 = np.minimum(sx + 1, w - 1)
        H = S[:, sx, :] * xt[:, 1][None, :, None] + S[:, x1, :] * xt[:, 2][None, :, None]
        r0, r1 = np.clip(yt[:, 0], 0, h - 1), np.clip(yt[:, 0] + 1, 0, h - 1)
        b0, b1 = yt[:, 1].astype(np.int64)[:, None, None], yt[:, 2].astype(np.int64)[:, None, None]
        return np.clip((((b0 * (H[r0] >> 4)) >> 16) + ((b1 * (H[r1] >> 4)) >> 16) + 2) >> 2, 0, 255).astype(np.uint8)

    rng = np.random.default_rng(1)
    shapes = [(int(rng.integers(1, 300)), int(rng.integers(1, 300))) for _ in range(30)] + [(448, 448), (224, 224), (1, 1), (15, 250)]
    for t, (h, w) in enumerate(shapes):
        oh, ow = (224, 224) if t % 3 else (int(rng.integers(1, 300)), int(rng.integers(1, 300)))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(img, (ow, oh)), model(img, oh, ow)), (h, w, oh, ow)
