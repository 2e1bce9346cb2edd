"""Known-answer test of the oracle's Philox4x32-10 (Random123 kat_vectors), the generator csrc/degrade.cu uses."""
import numpy as np

from oracle.degrade_oracle import philox4x32_10, philox_normals, philox_normals_v2

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_kat():
    for ctr, key, exp in KAT:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert tuple(int(x) for x in got) == exp


def test_normals_moments():
    z = philox_normals(seed=7, image_index=3, h=128, w=128).astype(np.float64)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs((z ** 3).mean()) < 0.05 and abs((z ** 4).mean() - 3.0) < 0.15
    # channels of one pixel and neighbouring pixels are uncorrelated
    assert abs(np.corrcoef(z[..., 0].ravel(), z[..., 1].ravel())[0, 1]) < 0.03
    assert abs(np.corrcoef(z[:, :-1, 0].ravel(), z[:, 1:, 0].ravel())[0, 1]) < 0.03
    # different images / seeds give different streams
    assert not np.array_equal(z, philox_normals(7, 4, 128, 128))
    assert not np.array_equal(z, philox_normals(8, 3, 128, 128))


def test_normals_v2_moments_and_layout():
    """noise stream 2 (four normals per Philox call, b2r_degrade since round 2)"""
    z = philox_normals_v2(seed=7, image_index=3, h=128, w=126).astype(np.float64)     # ragged width
    assert z.shape == (128, 126, 3)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs((z ** 3).mean()) < 0.05 and abs((z ** 4).mean() - 3.0) < 0.15
    assert abs(np.corrcoef(z[..., 0].ravel(), z[..., 1].ravel())[0, 1]) < 0.03
    assert abs(np.corrcoef(z[:, :-1, 0].ravel(), z[:, 1:, 0].ravel())[0, 1]) < 0.03
    assert abs(np.corrcoef(z[:-1, :, 2].ravel(), z[1:, :, 2].ravel())[0, 1]) < 0.03
    assert not np.array_equal(z, philox_normals_v2(7, 4, 128, 126))
    assert not np.array_equal(z, philox_normals_v2(8, 3, 128, 126))
    # every value of the image is drawn exactly once: no two positions share a normal
    assert len(np.unique(philox_normals_v2(7, 3, 16, 18).ravel())) == 16 * 18 * 3
