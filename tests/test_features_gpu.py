"""GPU parity for the feature half of the path: ORB detectAndCompute (bit-exact as a set), Hamming / L2 matchers
(bit-exact incl. order), RANSAC homography (< 0.5 px, in practice ~1e-6 px) -- all through the C ABI, against live cv2 /
the oracle restatements / the reference-generated goldens."""
import numpy as np
import cv2
import pytest
import torch

from oracle import orb as oorb, matching as omt, ransac as ors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import b200mosaic.ops as o
    return o


@pytest.fixture(scope="module")
def frames(golden_dir):
    return np.load(golden_dir / "clip01_frames.npz")["frames"]


def _synthetic_gray(w, h, seed):
    from b200mosaic.synth import DroneSweep
    return cv2.cvtColor(DroneSweep(w, h, seed=seed, ground_size=2048).next(), cv2.COLOR_BGR2GRAY)


def _orb_compare(ops, gray, nfeatures=700):
    kp, des = ops.orb_detect_and_compute(torch.from_numpy(gray).cuda(), nfeatures)
    kc, dc = oorb.cv_detect_and_compute(gray, nfeatures)
    a, ad = oorb.canon(kp.astype(np.float64), des)
    b, bd = oorb.canon(kc, dc)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a[:, [0, 1, 2, 5]], b[:, [0, 1, 2, 5]])          # pt, size, octave
    assert np.array_equal(a[:, 4], b[:, 4]), np.abs(a[:, 4] - b[:, 4]).max()   # Harris response bit-equal
    assert np.array_equal(a[:, 3], b[:, 3]), np.abs(a[:, 3] - b[:, 3]).max()   # IC angle bit-equal
    assert np.array_equal(ad, bd), np.mean((ad == bd).all(axis=1))
    # and the order is cv2's own (cvorder.cuh): row for row, no canonical sort needed
    assert np.array_equal(kp, kc.astype(np.float32)) and np.array_equal(des, dc)


def test_orb_pyramid_and_fast_scores_bit_exact(ops, frames):
    """INTER_LINEAR_EXACT resize chain and the FAST-9 score map of every level vs the oracle restatement."""
    g = cv2.cvtColor(frames[1], cv2.COLOR_BGR2GRAY)
    lev = oorb.build_pyramid(g)
    for l in range(8):
        img, sc = ops.orb_debug_level(torch.from_numpy(g).cuda(), l)
        assert np.array_equal(img, lev[l]), l
        assert np.array_equal(sc, oorb.fast_score_map(lev[l]).astype(np.uint8)), l
    assert np.array_equal(lev[1], cv2.resize(g, lev[1].shape[::-1], interpolation=cv2.INTER_LINEAR_EXACT))


@pytest.mark.parametrize("i", [0, 2, 4])
def test_orb_bit_exact_on_clip_frames(ops, frames, i):
    _orb_compare(ops, cv2.cvtColor(frames[i], cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("size", [(640, 360), (854, 480), (1280, 720), (1920, 1080)])
def test_orb_bit_exact_on_synthetic(ops, size):
    _orb_compare(ops, _synthetic_gray(size[0], size[1], 21))


@pytest.mark.parametrize("nfeatures,size", [(1000, (1280, 720)), (2000, (1280, 720)), (2000, (1920, 1080)), (300, (640, 360))])
def test_orb_other_feature_budgets(ops, nfeatures, size):
    """the reference's other ORB users run the same primitive with other budgets -- ORB_create(nfeatures=2000) in
    slam.py:47 (keyframes, slam.py:332) and ORB_create(nfeatures=1000) + BFMatcher(HAMMING, crossCheck=True) in depth_to_3d.py:856-889.
    Same kernels through the same stage entry points; the per-level quotas follow cv2's float32 geometric series."""
    g = _synthetic_gray(size[0], size[1], 33)
    _orb_compare(ops, g, nfeatures)
    if nfeatures == 1000:                                    # depth_to_3d.py:884-885: bf.match(prev_desc, curr_desc), sorted by distance
        g2 = np.roll(g, (3, 7), axis=(0, 1))
        _, d1 = ops.orb_detect_and_compute(torch.from_numpy(g).cuda(), nfeatures)
        _, d2 = ops.orb_detect_and_compute(torch.from_numpy(g2).cuda(), nfeatures)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
        cvm = sorted(bf.match(d1, d2), key=lambda m: m.distance)
        assert np.array_equal(ops.match_hamming_crosscheck(d1, d2), np.array([[m.queryIdx, m.trainIdx, m.distance] for m in cvm]))


def test_orb_featureless_image(ops):
    g = np.full((240, 320), 90, np.uint8)
    kp, des = ops.orb_detect_and_compute(torch.from_numpy(g).cuda())
    assert len(kp) == 0 and des.shape == (0, 32)


def test_hamming_crosscheck_exact(ops, golden_dir):
    g = np.load(golden_dir / "clip01_orb.npz")
    got = ops.match_hamming_crosscheck(g["des1"], g["des0"])
    assert np.array_equal(got, g["matches1"])
    # tie rules: duplicated rows on both sides
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8); t = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    q[10] = q[11]; t[5] = t[6]; q[20] = t[5]; t[100:110] = q[50:60]
    assert np.array_equal(ops.match_hamming_crosscheck(q, t), omt.match_hamming_crosscheck(q, t))
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    cvm = sorted(bf.match(q, t), key=lambda m: m.distance)
    assert np.array_equal(ops.match_hamming_crosscheck(q, t), np.array([[m.queryIdx, m.trainIdx, m.distance] for m in cvm]))


def test_l2_ratio_exact(ops, golden_dir):
    s = np.load(golden_dir / "clip01_sift.npz")
    got = ops.match_l2_ratio(s["des1"], s["des0"])
    assert np.array_equal(got, s["matches1"])
    assert np.array_equal(ops.match_l2_ratio(s["des0"], s["des1"]), omt.match_l2_ratio(s["des0"], s["des1"]))


@pytest.mark.parametrize("nq,nt", [(1, 2), (5, 3), (127, 255), (128, 256), (129, 257), (700, 701), (1500, 900), (300, 2)])
def test_l2_ratio_tensor_core_tiles_and_ties(ops, nq, nt):
    """tcgen05 matcher (match_tc.cu): query tiles of 128, train tiles of 256, ragged edges, exact duplicates (distance ties must
    go to the lower train index) and the k=2 corner cases -- bit-exact against the oracle's stable argsort."""
    rng = np.random.default_rng(nq * 1000 + nt)
    base = rng.integers(0, 60, (max(nt // 3, 1), 128))
    t = base[rng.integers(0, len(base), nt)] + (rng.random((nt, 128)) < 0.02) * rng.integers(0, 196, (nt, 128))
    q = t[rng.integers(0, nt, nq)] + (rng.random((nq, 128)) < 0.05) * rng.integers(0, 60, (nq, 128))
    q = np.clip(q, 0, 255).astype(np.float32); t = np.clip(t, 0, 255).astype(np.float32)
    for ratio in (0.7, 1.01):
        assert np.array_equal(ops.match_l2_ratio(q, t, ratio), omt.match_l2_ratio(q, t, ratio))


def _reproj(Ha, Hb, w, h):
    ys, xs = np.mgrid[0:h:16, 0:w:16]
    p = np.stack([xs.ravel(), ys.ravel(), np.ones(xs.size)])
    a = Ha @ p; b = Hb @ p
    return np.abs(a[:2] / a[2] - b[:2] / b[2]).max()


@pytest.mark.parametrize("det", ["orb", "sift"])
def test_ransac_on_golden_matches(ops, golden_dir, det):
    g = np.load(golden_dir / f"clip01_{det}.npz")
    mm = g["matches1"]
    src = g["kp1"][mm[:, 0].astype(int), :2].astype(np.float32)
    dst = g["kp0"][mm[:, 1].astype(int), :2].astype(np.float32)
    Hc, mask = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
    Ho, tr = ors.find_homography_ransac(src, dst, return_trace=True)
    H, iters, ninl = ops.ransac_homography(src, dst)
    assert H is not None
    assert iters == tr["iters"]                       # same seeded hypothesis sequence and stopping rule
    assert _reproj(H, Hc, 427, 240) < 1e-3            # budget 0.5 px
    assert _reproj(H, Ho, 427, 240) < 1e-3


@pytest.mark.parametrize("frac", [0.0, 0.3, 0.6, 0.85])
def test_ransac_with_outliers(ops, frac):
    rng = np.random.default_rng(3)
    n = 400
    src = (rng.random((n, 2)) * [1920, 1080]).astype(np.float32)
    Ht = np.array([[1.02, 0.03, 5], [-0.02, 0.98, -7], [1e-5, 2e-5, 1]])
    p = np.c_[src, np.ones(n)] @ Ht.T
    dst = (p[:, :2] / p[:, 2:] + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
    k = int(n * frac)
    dst[:k] = (rng.random((k, 2)) * [1920, 1080]).astype(np.float32)
    Hc, _ = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
    Ho, tr = ors.find_homography_ransac(src, dst, return_trace=True)
    H, iters, ninl = ops.ransac_homography(src, dst)
    assert H is not None and Hc is not None
    assert iters == tr["iters"], (iters, tr["iters"])
    assert _reproj(H, Hc, 1920, 1080) < 0.5


def test_ransac_degenerate_inputs(ops):
    src = np.array([[0, 0], [100, 0], [100, 100], [0, 100]], np.float32)
    dst = src * 1.5 + 7
    H, _, _ = ops.ransac_homography(src, dst)
    Hc, _ = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
    assert _reproj(H, Hc, 100, 100) < 1e-6
    H, _, _ = ops.ransac_homography(src[:3], dst[:3])
    assert H is None
    # all points collinear: the reference gets None
    line = np.stack([np.arange(20.0), 2 * np.arange(20.0)], 1).astype(np.float32)
    Hc, _ = cv2.findHomography(line.reshape(-1, 1, 2), (line + 3).reshape(-1, 1, 2), cv2.RANSAC, 2.0)
    H, _, _ = ops.ransac_homography(line, line + 3)
    assert (H is None) == (Hc is None)


def test_orb_process_frame_end_to_end(frames, golden_dir, capsys):
    """Drop-in run on the reference's own clip frames.  cv2's keypoint ORDER is whatever libstdc++'s nth_element leaves, ours
    is level/row-major, and RANSAC's seeded sampling depends on the order -- so the trajectory is checked stage by stage on
    the pipeline's own data: features == cv2's as a set, matches == the oracle matcher on those descriptors,
    H_rel == cv2.findHomography on those matches (< 0.5 px), control flow / smoothing / composition == RefMosaic's,
    canvas == the oracle's warp+blend driven with the same homographies (<= 1 LSB per step)."""
    import b200mosaic
    from oracle.mosaic_ref import RefMosaic, blend_step_cv
    g = np.load(golden_dir / "clip01_orb.npz")
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    assert np.array_equal(vm.output_img, g["canvas0"])
    def kparr(kps):
        return np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps])
    kp_prev, des_prev = kparr(vm.kp_prev), vm.des_prev
    kr, dr = oorb.canon(g["kp0"], g["des0"])
    a, ad = oorb.canon(kp_prev, des_prev)
    assert np.array_equal(a, kr) and np.array_equal(ad, dr)
    ref = RefMosaic(frames[0], detector_type="orb", float64_canvas=False)      # host-side control flow mirror
    canvas = g["canvas0"].copy()
    for t in range(1, len(frames)):
        before = vm.output_img.copy()
        vm.process_frame(frames[t], t)
        info = vm.last_info
        assert info.status == 0
        assert abs(info.n_matches - int(g["n_matches"][t - 1])) <= 3
        kp_cur, des_cur = kparr(vm.kp_prev), vm.des_prev                         # state advanced: prev == this frame
        kc, dc = oorb.canon(*oorb.cv_detect_and_compute(cv2.cvtColor(frames[t], cv2.COLOR_BGR2GRAY)))
        a, ad = oorb.canon(kp_cur, des_cur)
        assert np.array_equal(a, kc) and np.array_equal(ad, dc)
        mm = np.array([[m.queryIdx, m.trainIdx, m.distance] for m in vm.matches])
        assert np.array_equal(mm, omt.match_hamming_crosscheck(des_cur, des_prev))
        src = kp_cur[mm[:, 0].astype(int), :2].astype(np.float32); dst = kp_prev[mm[:, 1].astype(int), :2].astype(np.float32)
        Hc, _ = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
        H_rel = np.array(info.H_rel).reshape(3, 3)
        assert _reproj(H_rel, Hc, 427, 240) < 1e-3
        # host control flow on the same H_rel
        assert ref.validate_homography(H_rel)
        H_abs = ref.H_old @ ref.smooth_homography(H_rel)
        ref.H_old = H_abs
        assert np.abs(H_abs - vm.H).max() < 1e-9
        warped = cv2.warpPerspective(frames[t], vm.H, (before.shape[1], before.shape[0]), flags=cv2.INTER_LINEAR)
        want = blend_step_cv(before, warped)
        d = np.abs(want.astype(np.int16) - vm.output_img.astype(np.int16))
        assert d.max() <= 1 and np.mean(d > 0) < 2e-3
        kp_prev, des_prev = kp_cur, des_cur
    assert capsys.readouterr().out == ""    # no warnings were printed on this clip (as in the reference run)


def _lm_routes(lib):
    """(lm_iters, iterations solved by eigen-decomposition) of the last profile call"""
    import ctypes as C

    def run(src, dst):
        H = np.zeros(9); cyc = np.zeros(8, np.int64); lm = C.c_int(0); js = C.c_int(0)
        rc = lib.bm_ransac_profile(src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), len(src), 2.0, 2000, 0.995,
                                   H.ctypes.data_as(C.POINTER(C.c_double)), cyc.ctypes.data_as(C.c_void_p), C.byref(lm), C.byref(js))
        assert rc == 0
        return H.reshape(3, 3), lm.value, (js.value >> 8) & 0xFF
    return run


def test_ransac_ill_conditioned_polish(golden_dir):
    """Frame 359 of clip 01 (ORB): the inliers (398 of 549) cover only the right half of the frame and the normal matrix of the LM polish
    has condition ~1e15.  cv2 4.13 polishes ALL NINE elements of H (its callback asserts `J.cols == 9`) and solves the singular normal
    equations with a truncated eigen pseudo-inverse; on this set ten iterations end part-way down a flat valley (residual 197.14 against a
    floor of 193.45), 10 px away -- at the far frame corners -- from where an 8-parameter polish ends.  Pinned: the restatement and the
    device (both of its solve routes) follow cv2 to that point."""
    from b200mosaic import ops, _lib
    from oracle import ransac as orc
    lib = _lib.load()
    g = np.load(golden_dir / "ransac_illcond.npz")
    src, dst, Hcv = g["src"], g["dst"], g["H_cv"]
    Ho, tr = orc.find_homography_ransac(src, dst, 2.0, return_trace=True)
    assert _reproj(Ho, Hcv, 854, 480) < 5e-3                                  # restatement == cv2 (measured 1.4e-8 px with the restated cv2 Jacobi)
    prof = _lm_routes(lib)
    try:
        for force in (0, 1):
            assert lib.bm_debug_lm_force_eig(force) == 0
            H, iters, ninl = ops.ransac_homography(src, dst, 2.0)
            assert iters == tr["iters"] == 16 and ninl == int(tr["mask"].sum()) == 398
            err = _reproj(H, Hcv, 854, 480)
            Hp, lm_iters, n_eig = prof(src, dst)
            print(f"ill-conditioned polish, force_eig={force}: {err:.2e} px from cv2 at the frame corners, {lm_iters} LM iterations, {n_eig} by eigen-decomposition")
            assert err < 5e-3
            assert np.array_equal(Hp, H) and lm_iters == 10 and (n_eig == lm_iters if force else n_eig <= lm_iters)
    finally:
        lib.bm_debug_lm_force_eig(0)


@pytest.mark.parametrize("force", [0, 1])
def test_ransac_polish_on_partial_coverage(force):
    """consensus sets that cover the whole frame, a part of it, or one corner (ill-conditioned): the device polish (elimination route and
    forced eigen-decomposition route) against cv2.findHomography, compared where the points are (inside their bounding box +- 25 %)"""
    from b200mosaic import ops, _lib
    lib = _lib.load()
    prof = _lm_routes(lib)
    rng = np.random.default_rng(11)
    worst = 0.0
    n_eig_total = 0
    try:
        assert lib.bm_debug_lm_force_eig(force) == 0
        for case in range(36):
            n = int(rng.integers(20, 500))
            lo, hi = [((0, 0), (854, 480)), ((500, 100), (854, 300)), ((700, 0), (854, 60)), ((0, 0), (1920, 1080)), ((1500, 800), (1920, 1080)),
                      ((0, 0), (3840, 2160))][case % 6]
            P = rng.uniform(lo, hi, (n, 2))
            Ht = np.array([[1 + rng.normal(0, .01), rng.normal(0, .01), rng.normal(0, 8)], [rng.normal(0, .01), 1 + rng.normal(0, .01), rng.normal(0, 8)],
                           [rng.normal(0, 1e-6), rng.normal(0, 1e-6), 1]])
            q = np.c_[P, np.ones(n)] @ Ht.T
            Q = q[:, :2] / q[:, 2:] + rng.normal(0, 0.4, (n, 2))
            src = P.astype(np.float32); dst = Q.astype(np.float32)
            Hc, _ = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
            H, lm_iters, n_eig = prof(src, dst)
            assert Hc is not None and 1 <= lm_iters <= 10
            n_eig_total += n_eig
            assert n_eig == lm_iters if force else True
            ext = 0.25 * (np.array(hi) - np.array(lo))
            b0 = np.array(lo) - ext; b1 = np.array(hi) + ext
            box = np.array([[b0[0], b0[1], 1], [b1[0], b0[1], 1], [b1[0], b1[1], 1], [b0[0], b1[1], 1.0]]).T
            a = H @ box; b = Hc @ box
            e = float(np.abs(a[:2] / a[2] - b[:2] / b[2]).max())
            worst = max(worst, e)
            assert e < 0.02, (case, n, e)
    finally:
        lib.bm_debug_lm_force_eig(0)
    print(f"polish on partial coverage, force_eig={force}: worst {worst:.2e} px from cv2, {n_eig_total} eigen-decomposition iterations")
