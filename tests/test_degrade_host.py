"""Host-side tap construction (`degrade.motion_blur_kernel`) against the reference's OpenCV recipe: bit-identical
for every (degree, angle) the reference can draw (14_train_unified_advanced.py:54-55), via the committed fixture and,
where cv2 is importable, live."""
import numpy as np
import pytest

from _util import golden


def test_taps_match_golden_fixture_all_pairs():
    from b200restore import degrade
    taps = golden("blur_taps_ref.npz")["taps"]          # [14, 361, 15, 15] float32 from cv2 (make_golden.py)
    for d in range(2, 16):
        for a in range(0, 361):
            k = degrade.motion_blur_kernel(d, a).astype(np.float32)
            assert np.array_equal(k, taps[d - 2, a, :d, :d]), (d, a)


def test_taps_match_cv2_live_float64():
    cv2 = pytest.importorskip("cv2")
    from b200restore import degrade
    for d in (2, 5, 10, 12, 15):
        for a in (0, 1, 45, 90, 133, 180, 271, 359, 360):
            M = cv2.getRotationMatrix2D((d / 2, d / 2), a, 1)
            ref = cv2.warpAffine(np.diag(np.ones(d)), M, (d, d)) / d
            assert np.array_equal(degrade.motion_blur_kernel(d, a), ref), (d, a)


def test_known_properties_of_the_reference_kernel():
    """Facts recorded in SURVEY.md §7: the rotated line loses mass; angle 0/360 leaves the diagonal."""
    from b200restore import degrade
    assert abs(degrade.motion_blur_kernel(10, 45).sum() - 0.739) < 2e-3
    assert abs(degrade.motion_blur_kernel(12, 45).sum() - 0.714) < 2e-3
    k0 = degrade.motion_blur_kernel(7, 0)
    assert np.allclose(k0, np.eye(7) / 7)
    assert np.array_equal(degrade.motion_blur_kernel(7, 360), k0)


def test_param_builders():
    from b200restore import degrade, _lib as L
    p = degrade.compound_params(3)
    assert p.order == L.B2R_ORDER_BLUR_FOG_NOISE and (p.ksize == 10).all() and (p.fog_on == 1).all()
    assert np.allclose(p.fog_t, 0.5) and np.array_equal(p.fog_add, np.full(3, np.float32(0.9 * (1 - 0.5))))
    assert np.array_equal(p.sigma, np.full(3, np.float32(0.02 ** 0.5)))
    assert np.array_equal(p.taps[0, :100].reshape(10, 10), degrade.motion_blur_kernel(10, 45).astype(np.float32))
    q = degrade.random_params(2000, np.random.default_rng(0))
    assert q.order == L.B2R_ORDER_FOG_NOISE_BLUR
    for frac in ((q.fog_on == 1).mean(), (q.sigma > 0).mean(), (q.ksize > 0).mean()):
        assert 0.45 < frac < 0.55                                  # p = 0.5 each (14:26-28)
    ks = q.ksize[q.ksize > 0]
    assert ks.min() == 5 and ks.max() == 15                        # randint(5, 15) inclusive (14:54)
    sg = q.sigma[q.sigma > 0] ** 2
    assert 0.0099 < sg.min() and sg.max() < 0.0301                 # var U(0.01, 0.03) (14:47)
    t = q.fog_t[q.fog_on == 1]
    assert 1 - 0.7 * 1.2 - 1e-6 <= t.min() and t.max() <= 1 - 0.3 * 0.8 + 1e-6
    with pytest.raises(ValueError):
        degrade.motion_blur_kernel(16, 0)


def test_batch_builders_equal_the_per_image_setters():
    """compound_params / demo_params / blur_params / noise_params fill whole batches at once; the arrays must equal
    what the per-image setters (the form random_params uses) produce, and degree <= 1 must mean "no blur" (14:56)."""
    from b200restore import degrade as D
    n = 7
    ref = D.DegradeParams(n)
    for i in range(n):
        ref.set_blur(i, 10, 45)
        ref.set_fog(i, 0.5)
        ref.set_noise(i, 0.02)
    for got in (D.compound_params(n), D.demo_params(n)):
        for f in ("ksize", "taps", "fog_on", "fog_t", "fog_add", "sigma"):
            assert np.array_equal(getattr(got, f), getattr(ref, f)), f
    assert D.demo_params(n).order == 1 and D.demo_params(n).flags == 1 and D.compound_params(n).order == 0
    b = D.blur_params(n, 12, 30)
    k = D.motion_blur_kernel(12, 30).astype(np.float32).reshape(-1)
    assert (b.ksize == 12).all() and all(np.array_equal(b.taps[i, :144], k) for i in range(n)) and not b.taps[:, 144:].any()
    assert (D.blur_params(n, 1, 45).ksize == 0).all()
    z = D.noise_params(n, 0.03)
    assert np.allclose(z.sigma, np.float32(0.03 ** 0.5)) and z.flags == 1 and not z.fog_on.any()
    dev_flags = D.fog_params(n, np.random.default_rng(0))
    assert dev_flags.fog_on.all() and not (dev_flags.sigma > 0).any() and not (dev_flags.ksize > 1).any()
