"""Pin the generators oracle (02 / 03 / 04 / 13 / PSNR): against the committed outputs of the reference's own functions,
against those functions run live where /root/reference is mounted, and against cv2 / NumPy for the pieces the CUDA
kernels restate (the cv2.normalize arithmetic, the fog table, np.uint8 wrap-around)."""
import random

import cv2
import numpy as np
import pytest

from _util import NpShim, ReplayNormal, ReplayRandom, golden, have_reference, load_ref
from oracle import generators_oracle as GO


def test_script02_matches_golden_including_wraparound():
    g = golden("generators_ref.npz")
    wrapped = 0
    for i in range(len(g["images"])):
        noise = 0 + (float(g["var02"][i]) ** 0.5) * g["z"][i]
        out = GO.add_gaussian_noise_02(g["images"][i], noise)
        assert np.array_equal(out, g["out02"][i]), i
        x = g["images"][i] / 255 + noise
        wrapped += int(((x < 0) & (out > 128)).sum())
    assert wrapped > 100, "the fixture must exercise the negative -> mod 256 quirk"
    # the bright image with tiny variance takes the low_clip = 0 branch
    assert (g["images"][2] / 255 + (float(g["var02"][2]) ** 0.5) * g["z"][2]).min() >= 0


def test_script03_matches_golden_and_table_matches_cv2():
    g = golden("generators_ref.npz")
    for i, (d, a) in enumerate(g["cases03"]):
        assert np.array_equal(GO.apply_motion_blur_03(g["images"][i], int(d), float(a)), g["out03"][i]), (i, d, a)
    rng = np.random.default_rng(0)
    for _ in range(200):
        smin = int(rng.integers(0, 230))
        smax = int(rng.integers(smin + 1, 256))
        img = rng.integers(smin, smax + 1, size=(9, 11, 3), dtype=np.uint8)
        img[0, 0, 0], img[1, 1, 2] = smin, smax
        ref = img.copy()
        cv2.normalize(ref, ref, 0, 255, cv2.NORM_MINMAX)
        assert np.array_equal(GO.normalize_minmax_table(smin, smax)[img], ref), (smin, smax)
    const = np.full((4, 4, 3), 77, np.uint8)
    ref = const.copy()
    cv2.normalize(ref, ref, 0, 255, cv2.NORM_MINMAX)
    assert np.array_equal(GO.normalize_minmax_table(77, 77)[const], ref)


def test_script04_matches_golden():
    g = golden("generators_ref.npz")
    for i in range(len(g["images"])):
        out, t = GO.add_fog_04(g["images"][i], float(g["u04"][i]), float(g["inten04"][i]))
        assert np.array_equal(out, g["out04"][i]), i
        assert 0.1 <= t <= 0.9


def test_script13_distortions_match_golden():
    g = golden("generators_ref.npz")
    for i in range(len(g["images"])):
        b = GO.stress_add_blur(g["images"][i])
        f = GO.stress_add_fog(b)
        z = GO.stress_add_noise(f, 0 + (0.01 ** 0.5) * g["z"][i])
        assert np.array_equal(b, g["blur13"][i]) and np.array_equal(f, g["fog13"][i]) and np.array_equal(z, g["noise13"][i]), i


def test_host_fog_table_is_the_reference_expression():
    """The product's per-image fog table (generators.fog_table) against the oracle on all 256 byte values."""
    from b200restore import generators as G
    v = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    for u, inten in ((0.8, 0.8), (1.2, 0.8), (1.0, 0.5), (0.9, 1.0), (1.1, 0.1)):
        ref, t = GO.add_fog_04(v, u, inten)
        assert np.array_equal(G.fog_table(t)[v], ref)
    assert np.array_equal(G.fog_table(0.9)[v], GO.stress_add_fog(v))
    rng = random.Random(5)
    ts = [float(np.clip(1.0 - 0.8 * rng.uniform(0.8, 1.2), 0.1, 0.9)) for _ in range(3)]
    rng = random.Random(5)
    assert ts == [GO.add_fog_04(v, rng.uniform(0.8, 1.2))[1] for _ in range(3)]


def test_psnr_definition():
    a = np.zeros((4, 4, 3), np.uint8)
    b = a.copy()
    b[0, 0, 0] = 10
    assert GO.psnr_08(a, b) == pytest.approx(10 * np.log10(255 ** 2 / (100 / 48)))
    assert GO.psnr_08(a, a) == float("inf")


@pytest.mark.skipif(not have_reference(), reason="/root/reference not mounted")
def _ssim_exact_windows(a, b, data_range=255.0):
    """SSIM from EXACT integer window sums (the form csrc/ssim.cu evaluates): independent of scipy's float64 passes."""
    h, w, c = a.shape
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    total = 0.0
    for ch in range(c):
        x = a[..., ch].astype(np.int64)
        y = b[..., ch].astype(np.int64)

        def win(v):
            cs = np.pad(v.cumsum(0).cumsum(1), ((1, 0), (1, 0)))
            return cs[7:, 7:] - cs[:-7, 7:] - cs[7:, :-7] + cs[:-7, :-7]

        sx, sy, sxx, syy, sxy = win(x), win(y), win(x * x), win(y * y), win(x * y)
        a1 = (2 * sx * sy) / 2401.0 + C1
        b1 = (sx * sx + sy * sy) / 2401.0 + C1
        a2 = (2 * (49 * sxy - sx * sy)) / 2352.0 + C2
        b2 = ((49 * sxx - sx * sx) + (49 * syy - sy * sy)) / 2352.0 + C2
        total += float(((a1 * a2) / (b1 * b2)).mean())
    return total / c


def test_ssim_restatement_against_exact_integer_windows():
    """08:123 (skimage is absent: parity unpinned by execution).  The scipy-based restatement of skimage's algorithm
    must agree with an exact integer-window evaluation, give 1 for identical images, fall with distortion, be
    symmetric, and reject images smaller than the window as skimage does."""
    rng = np.random.default_rng(8)
    for (h, w) in ((7, 7), (9, 31), (64, 48), (224, 224)):
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
        s = GO.ssim_08(a, b)
        assert s == pytest.approx(_ssim_exact_windows(a, b), abs=1e-11), (h, w)
        assert GO.ssim_08(b, a) == pytest.approx(s, abs=1e-12)
        assert GO.ssim_08(a, a) == 1.0
        assert 0.0 < s < 1.0
    smooth = np.repeat(np.repeat(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8), 8, 0), 8, 1)
    s_lo = GO.ssim_08(smooth, np.clip(smooth.astype(int) + rng.integers(-40, 41, smooth.shape), 0, 255).astype(np.uint8))
    s_hi = GO.ssim_08(smooth, np.clip(smooth.astype(int) + rng.integers(-4, 5, smooth.shape), 0, 255).astype(np.uint8))
    assert s_lo < s_hi < 1.0
    assert GO.ssim_08(np.zeros((16, 16, 3), np.uint8), np.full((16, 16, 3), 255, np.uint8)) == pytest.approx(
        6.5025 / (255.0 ** 2 + 6.5025), rel=1e-12)          # constant images: S = C1 / (255^2 + C1)
    with pytest.raises(ValueError):
        GO.ssim_08(np.zeros((6, 20, 3), np.uint8), np.zeros((6, 20, 3), np.uint8))


def test_against_live_reference():
    r02, r03, r04 = load_ref("02_gen_noise.py"), load_ref("03_gen_blur.py"), load_ref("04_gen_fog.py")
    rng = np.random.default_rng(11)
    for trial in range(5):
        img = rng.integers(0, 256, (29, 35, 3), dtype=np.uint8)
        z = rng.standard_normal(img.shape)
        var = float(rng.uniform(0.005, 0.05))
        r02.np = NpShim(ReplayNormal(z))
        assert np.array_equal(r02.add_gaussian_noise(img, var=var), GO.add_gaussian_noise_02(img, 0 + var ** 0.5 * z))
        d, a = int(rng.integers(2, 16)), float(rng.integers(0, 361))
        assert np.array_equal(r03.apply_motion_blur(img, degree=d, angle=a), GO.apply_motion_blur_03(img, d, a))
        u = float(rng.uniform(0.8, 1.2))
        r04.random = ReplayRandom([u])
        assert np.array_equal(r04.add_fog(img, fog_intensity=0.8), GO.add_fog_04(img, u, 0.8)[0])
