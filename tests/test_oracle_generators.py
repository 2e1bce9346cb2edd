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
