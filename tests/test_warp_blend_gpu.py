"""GPU parity: ingest / warpPerspective / distanceTransform / GaussianBlur / blend through the C ABI vs the oracle
(cv2 4.13 with IPP off, i.e. exactly what the reference's warp() computes -- main.py:861-927)."""
import numpy as np
import cv2
import pytest
import torch

from oracle.mosaic_ref import RefMosaic, blend_step_cv
from oracle import cvmath as cm

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops():
    import b200mosaic.ops as o
    return o


@pytest.fixture(scope="module")
def rng():
    return np.random.default_rng(11)


@pytest.mark.parametrize("shape", [(64, 96), (97, 131), (480, 854), (1080, 1920)])
def test_gray_bit_exact(ops, rng, shape):
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    gray, bgrx = ops.ingest_bgr(dev(img))
    assert np.array_equal(gray.cpu().numpy(), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(bgrx.cpu().numpy()[..., :3], img)


def _rand_H(rng, tx, ty, persp=1e-5, rot=0.05):
    H = np.eye(3)
    H[:2, :2] += rng.normal(size=(2, 2)) * rot
    H[0, 2], H[1, 2] = tx, ty
    H[2, :2] = rng.normal(size=2) * persp
    return H


@pytest.mark.parametrize("case", range(8))
def test_warp_perspective_bit_exact(ops, rng, case):
    sh, sw = [(120, 160), (480, 854), (333, 517), (1080, 1920)][case % 4]
    img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    dh, dw = int(2 * sh), int(1.2 * sw)
    persp = [1e-6, 1e-5, 1e-4, 1e-3][case % 4] if case < 6 else 3e-3
    H = _rand_H(rng, (dw - sw) / 2 + rng.normal() * 20, dh - sh - 30 * (case + 1) % max(sh, 1), persp)
    if case == 7:
        H[0, 2] = -sw * 0.7                                   # mostly off-canvas
    ref = cv2.warpPerspective(img, H, (dw, dh), flags=cv2.INTER_LINEAR)
    got = ops.warp_perspective(dev(img), H, (dw, dh)).cpu().numpy()
    assert np.count_nonzero(ref != got) == 0


def test_warp_identity_and_empty(ops, rng):
    img = rng.integers(1, 256, (90, 130, 3), dtype=np.uint8)
    got = ops.warp_perspective(dev(img), np.eye(3), (130, 90)).cpu().numpy()
    assert np.array_equal(got, img)
    H = np.eye(3); H[0, 2] = 5000.0
    got = ops.warp_perspective(dev(img), H, (200, 200)).cpu().numpy()
    assert got.max() == 0


def _masks(rng):
    m1 = (rng.random((300, 420)) > 0.003).astype(np.uint8) * 255
    m1[80:220, 100:330] = 255
    m2 = np.full((700, 900), 255, np.uint8); m2[3, 5] = 0; m2[650, 880] = 0
    m3 = np.zeros((128, 160), np.uint8); m3[30:100, 40:130] = 255
    m4 = np.full((90, 2304), 255, np.uint8); m4[:, 0] = 0
    # odd sizes (no multiple of 4 / 8 / 16), a convex quad like a warped frame, rows longer than one 2048-pixel scan chunk,
    # distances of many 16-row blocks in every direction, and a mask without any zero pixel
    m5 = np.zeros((333, 517), np.uint8)
    cv2.fillConvexPoly(m5, np.array([[40, 20], [490, 55], [470, 310], [15, 280]], np.int32), 255)
    m5[150, 260] = 0
    m6 = np.full((70, 4500), 255, np.uint8); m6[33, 2047] = 0; m6[5, 4499] = 0; m6[60, 2048] = 0
    m7 = np.full((1301, 1203), 255, np.uint8); m7[0, 0] = 0; m7[1300, 1202] = 0; m7[640, 17] = 0
    m8 = np.full((100, 130), 255, np.uint8)
    m9 = (rng.random((257, 391)) > 0.0005).astype(np.uint8) * 255
    return [m1, m2, m3, m4, m5, m6, m7, m8, m9]


@pytest.mark.parametrize("i", range(9))
def test_distance_transform_bit_exact(ops, rng, i):
    m = _masks(rng)[i]
    ref = cv2.distanceTransform(m, cv2.DIST_L2, 3)
    got = ops.distance_transform(dev(m)).cpu().numpy()
    assert np.array_equal(ref, got), np.abs(ref - got).max()


def test_gaussian_blur31(ops, rng):
    w = rng.random((200, 320)).astype(np.float32)
    ref = cv2.GaussianBlur(w, (31, 31), 0)
    got = ops.gaussian_blur31(dev(w)).cpu().numpy()
    assert np.abs(ref - got).max() <= 2e-7
    assert np.mean(ref == got) > 0.99        # same FMA order as cv2's AVX2 path (scalar tails aside)


def _scene(rng, hc=400, wc=512, overlap=True):
    canvas = np.zeros((hc, wc, 3), np.uint8)
    canvas[hc // 2:hc - 10, 40:wc - 40] = rng.integers(0, 256, (hc - 10 - hc // 2, wc - 80, 3), dtype=np.uint8)
    canvas[hc // 2 + 30:hc // 2 + 35, 100:110] = 0
    frame = cv2.GaussianBlur(rng.integers(0, 256, (hc // 2, wc - 80, 3), dtype=np.uint8), (5, 5), 0)
    frame[10:14, 20:30] = 0
    H = np.array([[1.01, 0.02, 45.0], [-0.015, 0.99, hc * 0.3 if overlap else -hc], [1e-5, -2e-5, 1.0]])
    warped = cv2.warpPerspective(frame, H, (wc, hc), flags=cv2.INTER_LINEAR)
    return canvas, warped, frame, H


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("use_win", [False, True])
def test_blend_step_within_1lsb(ops, rng, overlap, use_win):
    canvas, warped, _, _ = _scene(rng, overlap=overlap)
    ref = blend_step_cv(canvas, warped)
    win = None
    if use_win:
        ys, xs = np.nonzero(warped.any(axis=2))
        win = (xs.min(), ys.min(), xs.max() + 1, ys.max() + 1) if len(xs) else (0, 0, 1, 1)
    got, any_ov = ops.blend_step(dev(canvas), dev(warped), win)
    got = got.cpu().numpy()
    assert any_ov == overlap
    d = np.abs(ref.astype(np.int16) - got.astype(np.int16))
    assert d.max() <= 1, (d.max(), np.argwhere(d > 1)[:5])
    assert np.mean(d > 0) < 2e-3


def test_handle_first_frame_and_warp_sequence(rng):
    """VideMosaic.__init__ paste + repeated warp(frame, H) on the device canvas vs the oracle, per step on identical
    (canvas_before, frame, H) and cumulatively."""
    import b200mosaic
    from b200mosaic.synth import DroneSweep
    sweep = DroneSweep(320, 200, seed=5, ground_size=1024, max_step=9.0)
    frames = sweep.frames(6)
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    ref = RefMosaic(frames[0], detector_type="orb", float64_canvas=False)
    assert (vm.h_offset, vm.w_offset) == (ref.h_offset, ref.w_offset)
    assert np.array_equal(vm.output_img, ref.output_img)
    H = ref.H_old.copy()
    worst = 0
    for t in range(1, 6):
        H = H @ sweep.D_true[t - 1]
        before = vm.output_img.copy()
        out = vm.warp(frames[t], H)
        # per-step parity on identical inputs
        warped = cv2.warpPerspective(frames[t], H, (before.shape[1], before.shape[0]), flags=cv2.INTER_LINEAR)
        want = blend_step_cv(before, warped)
        d = np.abs(want.astype(np.int16) - out.astype(np.int16))
        worst = max(worst, int(d.max()))
        assert d.max() <= 1 and np.mean(d > 0) < 2e-3
        assert vm.last_info.any_overlap == 1
    assert worst <= 1


def test_full_size_1080p_chain_and_4k_distance_transform(ops, rng):
    """the bench configuration itself: 1080p frames into the default 2160 x 2304 canvas (window ~ 1.1k x 1.9k, distances of
    several hundred pixels), per step against the oracle on identical inputs; and the distance transform of a 4K-frame-sized
    quad (two scan chunks per row, 140 row blocks)."""
    import b200mosaic
    from b200mosaic.synth import DroneSweep
    sweep = DroneSweep(1920, 1080, seed=3, ground_size=4096, max_step=12.0)
    frames = sweep.frames(4)
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    H = vm.H_old.copy()
    for t in range(1, 4):
        H = H @ sweep.D_true[t - 1]
        before = vm.output_img.copy()
        out = vm.warp(frames[t], H)
        warped = cv2.warpPerspective(frames[t], H, (before.shape[1], before.shape[0]), flags=cv2.INTER_LINEAR)
        want = blend_step_cv(before, warped)
        d = np.abs(want.astype(np.int16) - out.astype(np.int16))
        assert d.max() <= 1 and np.mean(d > 0) < 2e-3, (t, d.max(), np.mean(d > 0))
    m = np.zeros((2240, 3904), np.uint8)
    cv2.fillConvexPoly(m, np.array([[30, 40], [3870, 25], [3890, 2200], [12, 2215]], np.int32), 255)
    m[1100, 2047] = 0
    assert np.array_equal(cv2.distanceTransform(m, cv2.DIST_L2, 3), ops.distance_transform(dev(m)).cpu().numpy())
