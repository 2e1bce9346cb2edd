"""Pin the degradation oracle: (a) against the reference's own functions run here with their random draws replayed
(skipped where /root/reference is absent), (b) against the committed outputs of those functions."""
import numpy as np
import pytest

from _util import NpShim, ReplayNormal, ReplayRandom, golden, have_reference, load_ref
from oracle import degrade_oracle as O


def test_compound16_matches_golden():
    g = golden("degrade_ref.npz")
    for i in range(len(g["images"])):
        noise = 0 + (0.02 ** 0.5) * g["z"][i]
        assert np.array_equal(O.compound_16(g["images"][i], noise), g["out16"][i])


def test_random14_matches_golden():
    g = golden("degrade_ref.npz")
    for i, (fog_t, var, d, a) in enumerate(g["meta14"]):
        noise = None if np.isnan(var) else 0 + (var ** 0.5) * g["z"][i]
        out = O.random_14(g["images"][i], None if np.isnan(fog_t) else float(fog_t), noise,
                          int(d) if d > 0 else None, float(a))
        assert np.array_equal(out, g["out14"][i]), i


def test_u8_roundtrip_claim():
    """csrc/degrade.cu relies on u8 -> f32/255 -> *255 -> truncate being the identity (16:21-22, 14:62-64)."""
    v = np.arange(256, dtype=np.uint8)
    assert np.array_equal(O.quant_u8(v.astype(np.float32) / 255.0), v)
    assert np.array_equal(np.clip((v.astype(np.float32) / 255.0).astype(np.float64) * 255, 0, 255).astype(np.uint8), v)


@pytest.mark.skipif(not have_reference(), reason="/root/reference not mounted")
def test_against_live_reference():
    r16 = load_ref("16_gen_compound_data.py")
    r14 = load_ref("14_train_unified_advanced.py")
    rng = np.random.default_rng(42)
    for trial in range(6):
        img = rng.integers(0, 256, (33, 47, 3), dtype=np.uint8)
        z = rng.standard_normal(img.shape)
        r16.np = NpShim(ReplayNormal(z))
        assert np.array_equal(r16.apply_compound_distortion(img), O.compound_16(img, (0.02 ** 0.5) * z))
        fog = (rng.uniform(0.3, 0.7), rng.uniform(0.8, 1.2)) if trial % 2 == 0 else None
        var = rng.uniform(0.01, 0.03) if trial % 3 != 0 else None
        blur = (int(rng.integers(5, 16)), int(rng.integers(0, 361))) if trial != 4 else None
        draws = ([0.1, *fog] if fog else [0.9]) + ([0.1, var] if var else [0.9]) + ([0.1, *blur] if blur else [0.9])
        r14.random = ReplayRandom(draws)
        r14.np = NpShim(ReplayNormal(z))
        ref = r14.apply_random_distortions(img)
        got = O.random_14(img, (1.0 - fog[0] * fog[1]) if fog else None, (var ** 0.5) * z if var else None,
                          blur[0] if blur else None, blur[1] if blur else None)
        assert np.array_equal(ref, got), trial
