"""Letterbox preprocessing without a GPU: the numpy restatement of Pillow's 8-bit bicubic
resampler against the fixtures written from the reference's resize_image, against Pillow itself where it is installed,
and the library's host-side coefficient function against the restatement.  Bit-exact everywhere (integer arithmetic)."""
import os

import numpy as np
import pytest

from oracle import letterbox_oracle as LO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "letterbox.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_matches_reference_fixtures_bit_exact(gold):
    for i, (iw, ih, w, h) in enumerate(gold["cases"]):
        for lb in (0, 1):
            got = LO.resize_image_u8(gold[f"src{i}"], (int(w), int(h)), bool(lb))
            assert np.array_equal(got, gold[f"dst{i}_lb{lb}"]), (i, lb)


def test_oracle_matches_pillow_on_random_sizes():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(7)
    for _ in range(12):
        iw, ih = int(rng.integers(5, 400)), int(rng.integers(5, 400))
        w, h = int(rng.integers(8, 200)), int(rng.integers(8, 200))
        img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        nw, nh, _, _ = LO.letterbox_geometry(iw, ih, w, h, True)
        if nw < 1 or nh < 1:
            continue
        want = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
        assert np.array_equal(LO.resize_bicubic_u8(img, nw, nh), want), (iw, ih, nw, nh)


@pytest.mark.parametrize("a,b", [(97, 96), (640, 640), (500, 640), (1280, 640), (4000, 640), (20, 640), (7, 32), (333, 128)])
def test_library_coefficients_equal_the_restatement(a, b):
    from transparent_object_detection_b200.preprocess import resample_coeffs
    b1, k1, ks1 = resample_coeffs(a, b)
    b2, k2, ks2 = LO.precompute_coeffs(a, b)
    assert ks1 == ks2 and np.array_equal(b1, b2) and np.array_equal(k1, k2)
    assert (k1.sum(1) - (1 << 22)).__abs__().max() <= k1.shape[1]          # weights sum to one up to rounding


def test_letterbox_geometry_matches_reference_formula():
    from transparent_object_detection_b200.preprocess import letterbox_geometry
    for iw, ih, w, h in [(640, 480, 640, 640), (480, 640, 640, 640), (1000, 10, 64, 64), (33, 77, 128, 96)]:
        for lb in (True, False):
            assert letterbox_geometry(iw, ih, w, h, lb) == LO.letterbox_geometry(iw, ih, w, h, lb)
