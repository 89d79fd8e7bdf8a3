"""Pins the oracle's ORB / matcher / RANSAC restatements against live cv2 4.13 and the reference-generated goldens."""
import numpy as np
import cv2
import pytest

from oracle import orb, matching as mt, ransac as rs


@pytest.fixture(scope="module")
def frames(golden_dir):
    return np.load(golden_dir / "clip01_frames.npz")["frames"]


def test_orb_restatement_bit_exact_vs_cv2(frames):
    for f in frames[:2]:
        g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
        kp, des = orb.canon(*orb.detect_and_compute(g))
        kc, dc = orb.canon(*orb.cv_detect_and_compute(g))
        assert kp.shape == kc.shape and np.array_equal(kp, kc) and np.array_equal(des, dc)


def test_orb_restatement_vs_reference_golden(frames, golden_dir):
    g = np.load(golden_dir / "clip01_orb.npz")
    kp, des = orb.canon(*orb.detect_and_compute(cv2.cvtColor(frames[0], cv2.COLOR_BGR2GRAY)))
    kr, dr = orb.canon(g["kp0"], g["des0"])
    assert np.array_equal(kp, kr) and np.array_equal(des, dr)


def test_orb_quotas_and_sizes():
    assert orb.level_quotas(700) == [152, 127, 106, 88, 73, 61, 51, 42]
    assert orb.level_sizes(854, 480) == [(854, 480), (712, 400), (593, 333), (494, 278), (412, 231), (343, 193), (286, 161), (238, 134)]


def test_resize_linear_exact(frames):
    g = cv2.cvtColor(frames[1], cv2.COLOR_BGR2GRAY)
    for (w, h) in [(356, 200), (301, 177), (100, 57)]:
        assert np.array_equal(orb.resize_linear_exact(g, w, h), cv2.resize(g, (w, h), interpolation=cv2.INTER_LINEAR_EXACT))


def test_matchers_vs_reference_golden(golden_dir):
    g = np.load(golden_dir / "clip01_orb.npz")
    assert np.array_equal(mt.match_hamming_crosscheck(g["des1"], g["des0"]), g["matches1"])
    s = np.load(golden_dir / "clip01_sift.npz")
    assert np.array_equal(mt.match_l2_ratio(s["des1"], s["des0"]), s["matches1"])


def _reproj(Ha, Hb, w, h):
    ys, xs = np.mgrid[0:h:16, 0:w:16]
    p = np.stack([xs.ravel(), ys.ravel(), np.ones(xs.size)])
    a = Ha @ p; b = Hb @ p
    return np.abs(a[:2] / a[2] - b[:2] / b[2]).max()


@pytest.mark.parametrize("det", ["orb", "sift"])
def test_ransac_restatement_vs_cv2_on_golden_matches(golden_dir, det):
    g = np.load(golden_dir / f"clip01_{det}.npz")
    mm = g["matches1"]
    src = g["kp1"][mm[:, 0].astype(int), :2].astype(np.float32)
    dst = g["kp0"][mm[:, 1].astype(int), :2].astype(np.float32)
    Hc, _ = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
    H = rs.find_homography_ransac(src, dst)
    assert _reproj(H, Hc, 427, 240) < 1e-6


@pytest.mark.parametrize("frac", [0.3, 0.6])
def test_ransac_restatement_with_outliers(frac):
    rng = np.random.default_rng(3)
    n = 300
    src = (rng.random((n, 2)) * [640, 360]).astype(np.float32)
    Ht = np.array([[1.02, 0.03, 5], [-0.02, 0.98, -7], [1e-5, 2e-5, 1]])
    p = np.c_[src, np.ones(n)] @ Ht.T
    dst = (p[:, :2] / p[:, 2:] + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
    k = int(n * frac)
    dst[:k] = (rng.random((k, 2)) * [640, 360]).astype(np.float32)
    Hc, _ = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 2.0)
    H = rs.find_homography_ransac(src, dst)
    assert _reproj(H, Hc, 640, 360) < 1e-5


def test_lm_polish_is_cv2s_nine_parameter_form(golden_dir):
    """cv2 4.13's homography polish runs over all nine elements of H with truncated eigen pseudo-inverses (oracle/ransac.py::lm_refine).
    An 8-parameter polish converges to the same place on well-conditioned sets, so only an ill-conditioned one tells them apart: frame 359
    of clip 01 (ORB), whose 398 inliers cover the right half of the frame -- cv2's ten iterations stop at a residual of 197.14, the
    valley floor (where an 8-parameter polish ends) is 193.45 and lies 10 px away at the far frame corners."""
    g = np.load(golden_dir / "ransac_illcond.npz")
    src, dst, Hcv = g["src"], g["dst"], g["H_cv"]
    H, tr = rs.find_homography_ransac(src, dst, 2.0, return_trace=True)
    m = tr["mask"]
    assert tr["iters"] == 16 and int(m.sum()) == 398
    assert _reproj(H, Hcv, 854, 480) < 1e-5                     # measured 1.4e-8 px (9e-4 with LAPACK's eigh in place of cv2's Jacobi: cond 1e15)
    q = np.c_[src[m].astype(np.float64), np.ones(int(m.sum()))] @ H.T
    res = float(np.sum((q[:, :2] / q[:, 2:] - dst[m]) ** 2))
    assert abs(res - 197.136) < 0.01
    # cv2.findHomography(method=0) is runKernel + the same polish: the restatement follows it on sets that cover a corner of the frame
    rng = np.random.default_rng(1)
    for case in range(24):
        n = int(rng.integers(20, 400))
        lo, hi = [((0, 0), (854, 480)), ((500, 100), (854, 300)), ((700, 0), (854, 60))][case % 3]
        P = rng.uniform(lo, hi, (n, 2))
        Ht = np.array([[1 + rng.normal(0, .01), rng.normal(0, .01), rng.normal(0, 8)], [rng.normal(0, .01), 1 + rng.normal(0, .01), rng.normal(0, 8)],
                       [rng.normal(0, 1e-5), rng.normal(0, 1e-5), 1]])
        q = np.c_[P, np.ones(n)] @ Ht.T
        Q = q[:, :2] / q[:, 2:] + rng.normal(0, 0.5, (n, 2))
        P32 = P.astype(np.float32); Q32 = Q.astype(np.float32)
        Hc, _ = cv2.findHomography(P32, Q32, 0)
        Hm = rs.lm_refine(rs.run_kernel(P32, Q32), P32, Q32)
        box = np.array([[lo[0], lo[1], 1], [hi[0], lo[1], 1], [hi[0], hi[1], 1], [lo[0], hi[1], 1.0]]).T
        a = Hm @ box; b = Hc @ box
        assert np.abs(a[:2] / a[2] - b[:2] / b[2]).max() < 1e-3, case


def test_cv_jacobi_is_cv2_eigen():
    """oracle.ransac.cv_jacobi restates the Jacobi inside cv::eigen / cv::solve(DECOMP_EIG): same eigenvalues and row eigenvectors as
    cv2.eigen to a few ulps, including the tiny eigenvalues of badly scaled matrices (where the pseudo-inverse's truncation rule looks)"""
    rng = np.random.default_rng(0)
    for case in range(8):
        n = [2, 3, 5, 9, 9, 9, 9, 9][case]
        B = rng.normal(size=(n, n)); A = B @ B.T
        if case >= 4:
            d = np.diag(10.0 ** rng.uniform(-3, 6, n)); A = d @ A @ d
        ok, w, v = cv2.eigen(A)
        W, V = rs.cv_jacobi(A)
        assert np.all(np.abs(W - w.ravel()) <= 1e-11 * np.abs(w.ravel()) + 1e-300)     # RELATIVE, eigenvalue by eigenvalue (measured <= 4e-13)
        assert np.abs(V - v).max() < 1e-12


def test_rng_first_draws():
    r = rs.CvRNG()
    v = [r.next() for _ in range(3)]
    c = cv2.RNG if hasattr(cv2, "RNG") else None
    assert v[0] == ((0xFFFFFFFF * 4164903690 + 0xFFFFFFFF) & 0xFFFFFFFF)
