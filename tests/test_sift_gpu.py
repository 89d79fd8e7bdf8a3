"""GPU parity for SIFT (cv2.SIFT_create(700), main.py:33,112,718).  SIFT is a float32 pipeline whose near-threshold
decisions depend on the last bits of the pyramid, and cv2's own SIFT is not bit-repeatable from call to call (orientation angles
jitter by an ulp, a keypoint at the 0.8-of-maximum orientation threshold may come or go: tests/test_oracle_order_cpu.py).  Stated
tolerances (north_star: "SIFT descriptors within a stated L2 tolerance"), all as MAXIMA over the reproduced keypoints:
  * Gaussian / DoG pyramid: max |diff| <= 2e-4 grey levels vs the cv2 primitive chain (same FMA order as cv2's AVX2 filters)
  * keypoints: >= 99.5 % of cv2's keypoints reproduced at |dx|+|dy|+|dsize| < 0.02 px (measured: 100 % on every frame below),
    EVERY one of them with |dangle| < 0.01 deg, response within 1e-6 (8 float32 ulps) and the packed octave / layer field exact
  * descriptors: EVERY reproduced keypoint's descriptor within L2 16 of cv2's (3 % of the 512 norm; measured: max 1, >= 99.8 % exact)
"""
import numpy as np
import cv2
import pytest
import torch

from oracle import sift as osift

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import b200mosaic.ops as o
    return o


@pytest.fixture(scope="module")
def frames(golden_dir):
    return np.load(golden_dir / "clip01_frames.npz")["frames"]


def test_sift_pyramid_close_to_cv2_chain(ops, frames):
    g = cv2.cvtColor(frames[0], cv2.COLOR_BGR2GRAY)
    gp, dp = osift.build_pyramids(g)
    dev = torch.from_numpy(g).cuda()
    _, noct = ops.sift_debug_level(dev, 0, 0)
    assert noct == len(gp)
    worst = 0.0
    # octaves smaller than the largest kernel (27 taps) are skipped: cv2's border handling there is multi-bounce and those
    # octaves cannot hold keypoints anyway (5 px image border).  Bit-exactness vs cv2 depends on the HOST cpu's SIMD
    # dispatch (AVX2 vs AVX-512 accumulate in different orders), hence a tolerance of a few float32 ulps at 255.
    octs = [o for o in range(noct) if min(gp[o][0].shape) > 27]
    for o in octs:
        for l in range(6):
            img, _ = ops.sift_debug_level(dev, o, l)
            assert img.shape == gp[o][l].shape
            worst = max(worst, float(np.abs(img - gp[o][l]).max()))
        for l in range(5):
            d, _ = ops.sift_debug_level(dev, o, l, dog=True)
            worst = max(worst, float(np.abs(d - dp[o][l]).max()))
    assert worst <= 2e-4, worst


def _compare(ops, gray, min_frac=0.995):
    kp, des = ops.sift_detect_and_compute(torch.from_numpy(gray).cuda())
    kc, dc = osift.cv_detect_and_compute(gray)
    assert abs(len(kp) - len(kc)) <= 0.005 * len(kc) + 2, (len(kp), len(kc))
    assert np.array_equal(des, np.floor(des)) and des.min() >= 0 and des.max() <= 255
    pairs = osift.match_keypoints(kc, kp.astype(np.float64))
    frac = len(pairs) / max(len(kc), 1)
    assert frac >= min_frac, frac
    ia, ib = pairs[:, 0], pairs[:, 1]
    dang = np.abs(((kp[ib, 3] - kc[ia, 3]) + 180.0) % 360.0 - 180.0)
    assert dang.max() < 0.01, dang.max()                                                       # every reproduced keypoint, not a percentile
    assert np.array_equal(kp[ib, 5].astype(np.int64), kc[ia, 5].astype(np.int64))              # packed octave/layer/xi
    assert np.abs(kp[ib, 4] - kc[ia, 4]).max() < 1e-6                                        # response
    l2 = np.linalg.norm(des[ib].astype(np.float64) - dc[ia].astype(np.float64), axis=1)
    assert l2.max() <= 16.0, (l2.max(), int((l2 > 16).sum()))                                  # a maximum, not a percentile
    assert np.mean(l2 == 0) >= 0.98, np.mean(l2 == 0)
    return frac


@pytest.mark.parametrize("i", [0, 3])
def test_sift_on_clip_frames(ops, frames, i):
    _compare(ops, cv2.cvtColor(frames[i], cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("size,seed", [((640, 360), 9), ((1920, 1080), 9), ((1920, 1080), 31), ((1280, 720), 5)])
def test_sift_on_synthetic(ops, size, seed):
    from b200mosaic.synth import DroneSweep
    g = cv2.cvtColor(DroneSweep(size[0], size[1], seed=seed, ground_size=2048).next(), cv2.COLOR_BGR2GRAY)
    _compare(ops, g)


def test_sift_on_4k_frame(ops):
    """config 5 frame size (3840 x 2160): ~220 k refined candidates, 4x the 1080p lists -- list capacities scale with the frame
    (an overflow would surface as BM_ERR_UNSUPPORTED, never as silently truncated lists)"""
    from b200mosaic.synth import DroneSweep, make_ground
    ground = cv2.resize(make_ground(4096, 77), (8192, 8192))
    g = cv2.cvtColor(DroneSweep(3840, 2160, seed=77, ground=ground, max_step=40.0).next(), cv2.COLOR_BGR2GRAY)
    _compare(ops, g)


def test_sift_full_clip_frames(ops, golden_dir):
    """every 37th frame of the real clip at its native 854 x 480"""
    cap = cv2.VideoCapture(str(golden_dir / "clip01.mp4"))
    t, n = 0, 0
    while True:
        ok, f = cap.read()
        if not ok:
            break
        if t % 37 == 0:
            _compare(ops, cv2.cvtColor(f, cv2.COLOR_BGR2GRAY))
            n += 1
        t += 1
    assert n >= 15


def test_sift_output_order_of_rich_frames_is_cv2_keypoint_lessthan(ops):
    """frames with more than 12 288 candidates (the 1080p synthetic sweep: ~55 k) keep KeyPoint_LessThan order -- cv2's retainBest order
    would need an orientation histogram for every candidate there; smaller frames come out in cv2's own order (tests/test_order_gpu.py)"""
    from b200mosaic.synth import DroneSweep
    g = cv2.cvtColor(DroneSweep(1920, 1080, seed=9, ground_size=2048).next(), cv2.COLOR_BGR2GRAY)
    kp, _ = ops.sift_detect_and_compute(torch.from_numpy(g).cuda())
    key = [(r[0], r[1], -r[2], r[3], -r[4], -r[5]) for r in kp]
    assert key == sorted(key)


def test_sift_process_frame_end_to_end(frames):
    """stage-by-stage on the pipeline's own data (see the ORB end-to-end test for why): matches == oracle matcher on the
    device descriptors, H_rel == cv2.findHomography on those matches, canvas == oracle blend with the same H."""
    import b200mosaic
    from oracle import matching as omt
    from oracle.mosaic_ref import blend_step_cv
    vm = b200mosaic.VideMosaic(frames[0], detector_type="sift", show_intermediate=False, visualize=False)
    def kparr(kps):
        return np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps])
    kp_prev, des_prev = kparr(vm.kp_prev), vm.des_prev
    for t in range(1, 4):
        before = vm.output_img.copy()
        vm.process_frame(frames[t], t)
        info = vm.last_info
        assert info.status == 0 and info.n_matches > 150
        kp_cur, des_cur = kparr(vm.kp_prev), vm.des_prev
        mm = np.array([[m.queryIdx, m.trainIdx, m.distance] for m in vm.matches])
        assert np.array_equal(mm, omt.match_l2_ratio(des_cur, des_prev))
        src = kp_cur[mm[:, 0].astype(int), :2].astype(np.float32); dst = kp_prev[mm[:, 1].astype(int), :2].astype(np.float32)
        Hc, _ = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
        H_rel = np.array(info.H_rel).reshape(3, 3)
        ys, xs = np.mgrid[0:240:16, 0:427:16]
        p = np.stack([xs.ravel(), ys.ravel(), np.ones(xs.size)])
        a = H_rel @ p; b = Hc @ p
        assert np.abs(a[:2] / a[2] - b[:2] / b[2]).max() < 1e-3
        warped = cv2.warpPerspective(frames[t], vm.H, (before.shape[1], before.shape[0]), flags=cv2.INTER_LINEAR)
        d = np.abs(blend_step_cv(before, warped).astype(np.int16) - vm.output_img.astype(np.int16))
        assert d.max() <= 1
        kp_prev, des_prev = kp_cur, des_cur
