"""Host-side unit test of csrc/eig9.h -- the eigen-decomposition / truncated pseudo-inverse the homography polish (k_ransac_refine)
falls back to on ill-conditioned consensus sets.  The header is plain C++ (host + device), compiled here with g++ and compared with
what cv2 itself computes: cv2.eigen (the Jacobi inside cv::solve / cv::invert with DECOMP_EIG), cv2.solve and cv2.invert."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

CSRC = Path(__file__).resolve().parent.parent / "real-time-video-mosaic_b200" / "csrc"
WRAP = r'''
#include "eig9.h"
extern "C" int eig9(const double* A, const double* b, double* w, double* v, double* x, double* dp, int* kept) {
    double a[81];
    for (int i = 0; i < 81; ++i) a[i] = A[i];
    const int sweeps = bm_jacobi9(a, w, v);
    *kept = bm_eig_pinv9(w, v, b, x, dp);
    return sweeps;
}
'''


@pytest.fixture(scope="module")
def eig9(tmp_path_factory):
    d = tmp_path_factory.mktemp("eig9")
    (d / "wrap.cpp").write_text(WRAP)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", f"-I{CSRC}", "-o", str(d / "libeig9.so"), str(d / "wrap.cpp")])
    lib = C.CDLL(str(d / "libeig9.so"))

    def run(A, b):
        A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
        w = np.zeros(9); v = np.zeros((9, 9)); x = np.zeros(9); dp = np.zeros(9); kept = C.c_int(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        sweeps = lib.eig9(p(A), p(b), p(w), p(v), p(x), p(dp), C.byref(kept))
        return w, v, x, dp, kept.value, sweeps
    return run


def _normal_matrix(rng, box, n=300, noise=1.0):
    """J^T J, J^T r of the 9-parameter homography residuals on points inside `box` (what the polish builds)"""
    P = rng.uniform(box[0], box[1], (n, 2))
    h = np.array([1.001, 0.002, 3.0, -0.001, 0.998, -2.0, 1e-6, -2e-6, 1.0])
    ww = 1.0 / (h[6] * P[:, 0] + h[7] * P[:, 1] + h[8])
    xi = (h[0] * P[:, 0] + h[1] * P[:, 1] + h[2]) * ww; yi = (h[3] * P[:, 0] + h[4] * P[:, 1] + h[5]) * ww
    J = np.zeros((2 * n, 9))
    J[0::2, 0] = P[:, 0] * ww; J[0::2, 1] = P[:, 1] * ww; J[0::2, 2] = ww; J[0::2, 6] = -P[:, 0] * ww * xi; J[0::2, 7] = -P[:, 1] * ww * xi; J[0::2, 8] = -ww * xi
    J[1::2, 3] = P[:, 0] * ww; J[1::2, 4] = P[:, 1] * ww; J[1::2, 5] = ww; J[1::2, 6] = -P[:, 0] * ww * yi; J[1::2, 7] = -P[:, 1] * ww * yi; J[1::2, 8] = -ww * yi
    r = rng.normal(0, noise, 2 * n)
    return J.T @ J, J.T @ r


def test_jacobi9_matches_cv2_eigen_and_solve(eig9):
    import cv2
    rng = np.random.default_rng(5)
    boxes = [((0, 0), (854, 480)), ((0, 0), (1920, 1080)), ((500, 100), (854, 300)), ((700, 0), (854, 60)), ((0, 0), (3840, 2160))]
    for case in range(40):
        A, b = _normal_matrix(rng, boxes[case % len(boxes)])
        if case % 2:
            A = A + np.diag(0.3 * np.diag(A))                      # a damped system (lambda > 0): nothing is truncated
        w, v, x, dp, kept, sweeps = eig9(A, b)
        assert sweeps < 40
        # decomposition: orthonormal vectors, A v = w v to the accuracy of the data
        assert np.abs(v.T @ v - np.eye(9)).max() < 1e-13
        assert np.abs(A @ v - v * w).max() <= 1e-13 * np.abs(A).max()
        ok, wc, vc = cv2.eigen(A)
        wc = wc.ravel()
        ws = np.sort(w)[::-1]
        big = np.abs(wc) > 1e-12 * np.abs(wc).max()                 # the gauge eigenvalue of the undamped system is rounding noise
        assert np.allclose(ws[big], wc[big], rtol=1e-9, atol=0)
        # truncated pseudo-inverse: same eigenvalues kept, same solution and diagonal as cv2's own solve / invert
        thr = 2 * np.finfo(float).eps * wc.sum()
        assert kept == int((np.abs(wc) > thr).sum())
        xc = cv2.solve(A, b.reshape(-1, 1), flags=cv2.DECOMP_EIG)[1].ravel()
        dc = np.diag(cv2.invert(A, flags=cv2.DECOMP_EIG)[1])
        assert np.abs(x - xc).max() <= 1e-6 * np.abs(xc).max()
        assert np.abs(dp - dc).max() <= 1e-6 * np.abs(dc).max()
        if case % 2 == 0:
            assert kept <= 8                                         # the scale gauge is always dropped


def test_jacobi9_degenerate_inputs(eig9):
    w, v, x, dp, kept, sweeps = eig9(np.zeros((9, 9)), np.ones(9))
    assert kept == 0 and not x.any() and not dp.any()
    w, v, x, dp, kept, sweeps = eig9(np.diag(np.arange(1.0, 10.0)), np.ones(9))
    assert kept == 9 and np.allclose(x, 1.0 / np.arange(1.0, 10.0)) and sweeps == 0
    A = np.diag([1.0] * 8 + [2.0 * np.finfo(float).eps])             # at cv2's threshold: |w| <= 2 eps sum(w) is dropped
    w, v, x, dp, kept, sweeps = eig9(A, np.ones(9))
    assert kept == 8 and x[8] == 0.0
