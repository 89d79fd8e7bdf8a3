"""Pins oracle/finalize.py: (1) the call-for-call restatement against the UNMODIFIED reference functions (fixture
tests/golden/finalize.npz), (2) the arithmetic restatement (crop rectangle, cv2.resize INTER_LINEAR on 8-bit) bit-exact
against live cv2."""
import hashlib

import numpy as np
import cv2
import pytest

from oracle import finalize as fin


@pytest.fixture(scope="module")
def canvas(golden_dir):
    return np.load(golden_dir / "clip01_orb.npz")["canvas_final"]


@pytest.mark.parametrize("name,thr,margin", [("main", 80, 30), ("default", 15, 5)])
def test_restatement_matches_reference_fixture(golden_dir, canvas, name, thr, margin):
    g = np.load(golden_dir / "finalize.npz")
    crop = fin.crop_black_areas(canvas, threshold=thr, margin=margin)
    assert list(crop.shape) == list(g[f"{name}_crop_shape"]) and np.array_equal(crop[0, 0], g[f"{name}_crop_first"])
    scaled = fin.scale_to_screen(crop)
    assert list(scaled.shape) == list(g[f"{name}_scaled_shape"])
    assert np.array_equal(scaled[::8, ::8], g[f"{name}_scaled_sub8"])
    assert hashlib.sha256(np.ascontiguousarray(scaled).tobytes()).digest() == g[f"{name}_scaled_sha256"].tobytes()
    # the arithmetic layer gives the same image
    got, rect = fin.finalize(canvas, thr, margin)
    assert np.array_equal(got, scaled)
    assert np.array_equal(fin.scale_to_screen(canvas, 320, 300), g["small_scaled"])
    assert np.array_equal(fin.finalize(canvas, 255, 0, 320, 300)[0], g["small_scaled"])      # nothing above threshold -> whole canvas


@pytest.mark.parametrize("shape,dsize", [((300, 400), (1920, 1440)), ((517, 333), (640, 993)), ((1000, 1200), (777, 648)),
                                         ((64, 64), (1920, 1920)), ((480, 640), (320, 240)), ((97, 131), (131, 97)), ((5, 7), (40, 3))])
def test_resize_linear_bit_exact(shape, dsize):
    rng = np.random.default_rng(shape[0] * 7 + dsize[0])
    src = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    assert np.array_equal(fin.resize_linear_u8(src, *dsize), cv2.resize(src, dsize, interpolation=cv2.INTER_LINEAR))


def test_crop_rect_matches_cv2(canvas):
    for thr, margin in [(80, 30), (15, 5), (0, 0), (200, 100)]:
        want = fin.crop_black_areas(canvas, thr, margin)
        r = fin.crop_rect(canvas, thr, margin)
        x, y, w, h = r
        assert np.array_equal(canvas[y:y + h, x:x + w], want)
    assert fin.crop_rect(np.zeros((20, 30, 3), np.uint8), 80, 30) is None
