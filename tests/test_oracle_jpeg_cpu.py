"""oracle/jpeg.py (the restated libjpeg baseline encoder) against live cv2.imencode: the whole file, byte for byte."""
import numpy as np
import cv2
import pytest

from oracle import jpeg


def _img(h, w, kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "extreme":                                # saturated checkerboards: the largest coefficients the DCT can produce
        yy, xx = np.mgrid[0:h, 0:w]
        return (((yy // 3 + xx // 5) & 1) * 255).astype(np.uint8)[..., None].repeat(3, 2) ^ np.array([0, 255, 0], np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 3 + yy) % 256, (yy * 2 + xx // 2) % 256, (xx + yy * 5) % 256], -1).astype(np.uint8)
    return cv2.GaussianBlur(img, (0, 0), 2.0)


@pytest.mark.parametrize("size", [(1, 1), (8, 8), (16, 16), (17, 23), (100, 75), (121, 200), (240, 427), (30, 1), (1, 40)])
@pytest.mark.parametrize("kind", ["smooth", "noise", "extreme"])
def test_oracle_jpeg_equals_cv2_imencode(size, kind):
    img = _img(size[0], size[1], kind)
    ok, ref = cv2.imencode(".jpg", img)
    assert ok and jpeg.encode(img) == ref.tobytes()


@pytest.mark.parametrize("quality", [1, 30, 50, 75, 90, 100])
def test_oracle_jpeg_other_qualities(quality):
    img = _img(64, 80, "smooth", seed=quality)
    ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok and jpeg.encode(img, quality) == ref.tobytes()


def test_oracle_jpeg_on_a_mosaic_like_image():
    """the kind of image main.py:1664 writes: photographic content with a black rim (crop margin) -- decoded result also matches"""
    rng = np.random.default_rng(5)
    img = cv2.GaussianBlur(rng.integers(0, 256, (270, 480, 3), dtype=np.uint8), (0, 0), 1.2)
    img[:20] = 0
    img[:, -37:] = 0
    ok, ref = cv2.imencode(".jpg", img)
    mine = jpeg.encode(img)
    assert mine == ref.tobytes()
    assert np.array_equal(cv2.imdecode(np.frombuffer(mine, np.uint8), cv2.IMREAD_COLOR), cv2.imdecode(ref, cv2.IMREAD_COLOR))
