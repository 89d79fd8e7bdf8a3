"""Pins the oracle's restatements of the OpenCV primitives (oracle/cvmath.py) against live cv2 4.13 (IPP off)."""
import numpy as np
import cv2
import pytest

from oracle import cvmath as cm
from oracle.mosaic_ref import blend_step_cv


@pytest.fixture(scope="module")
def rng():
    return np.random.default_rng(7)


def test_gray_exact(rng):
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(cm.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("persp", [0.0, 1e-4, 2e-3])
def test_warp_perspective_exact(rng, persp):
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    H = np.eye(3)
    H[:2, :2] += rng.normal(size=(2, 2)) * 0.05
    H[0, 2], H[1, 2] = 20.3, 90.7
    H[2, :2] = [persp, -persp / 2]
    ref = cv2.warpPerspective(img, H, (200, 260), flags=cv2.INTER_LINEAR)
    assert np.array_equal(cm.warp_perspective(img, H, (200, 260)), ref)


def test_chamfer_dt_exact(rng):
    m = (rng.random((150, 210)) > 0.01).astype(np.uint8) * 255
    m[40:110, 50:170] = 255
    ref = cv2.distanceTransform(m, cv2.DIST_L2, 3)
    assert np.array_equal(cm.chamfer_dt(m), ref)
    assert np.array_equal(cm.chamfer_dt_closed_form(m), cm.chamfer_dt_int(m))


def test_chamfer_dt_large_distances():
    m = np.full((500, 640), 255, np.uint8)
    m[2, 3] = 0
    ref = cv2.distanceTransform(m, cv2.DIST_L2, 3)
    assert np.array_equal(cm.chamfer_dt(m), ref)


def test_blur31_matches_cv2(rng):
    w = rng.random((96, 128)).astype(np.float32)       # width % 8 == 0: no scalar tail in cv2's column filter
    ref = cv2.GaussianBlur(w, (31, 31), 0)
    k = cv2.getGaussianKernel(31, 5.0, cv2.CV_32F).ravel()
    assert np.array_equal(k, cm.gaussian_kernel_f32(31, 5.0))
    got = cm.blur31(w)
    assert np.abs(got - ref).max() <= 2e-7
    assert np.mean(got == ref) > 0.99


def _scene(rng, overlap=True):
    canvas = np.zeros((200, 256, 3), np.uint8)
    canvas[90:190, 40:220] = rng.integers(0, 256, (100, 180, 3), dtype=np.uint8)
    canvas[120:125, 100:110] = 0                                      # a hole of pure black inside the mosaic
    frame = rng.integers(0, 256, (100, 180, 3), dtype=np.uint8)
    frame[10:14, 20:30] = 0
    H = np.array([[1.01, 0.02, 45.0], [-0.015, 0.99, 60.0 if overlap else -200.0], [1e-5, -2e-5, 1.0]])
    warped = cv2.warpPerspective(frame, H, (256, 200), flags=cv2.INTER_LINEAR)
    return canvas, warped


@pytest.mark.parametrize("overlap", [True, False])
def test_blend_step_restatement_within_1lsb(rng, overlap):
    canvas, warped = _scene(rng, overlap)
    ref = blend_step_cv(canvas, warped)
    got = cm.blend_step(canvas, warped)
    d = np.abs(ref.astype(np.int16) - got.astype(np.int16))
    assert d.max() <= 1
    assert np.mean(d > 0) < 1e-3


def test_blend_step_float64_canvas_equals_uint8_canvas(rng):
    canvas, warped = _scene(rng)
    a = blend_step_cv(canvas.astype(np.float64), warped)
    b = blend_step_cv(canvas, warped)
    assert a.dtype == np.float64 and np.array_equal(a.astype(np.uint8), b) and np.array_equal(a, np.floor(a))
