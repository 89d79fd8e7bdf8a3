"""Restatement of cv2.findHomography(p_cur, p_prev, cv2.RANSAC, 2.0) as the reference calls it (main.py:856-857)
(TEST INFRASTRUCTURE, see oracle/__init__.py).  OpenCV 4.x calib3d: RANSACPointSetRegistrator::run +
HomographyEstimatorCallback + LMSolver refinement, from the published sources as pinned in SURVEY.md A.7 -- except the LM polish, which
in cv2 4.13 runs over all nine elements of H with truncated eigen pseudo-inverses (established from the installed binary and by probing
cv2.findHomography(method=0); see lm_refine and DESIGN.md section 2.1); checked against live cv2 4.13 in tests/test_oracle_features_cpu.py."""
from __future__ import annotations

import math

import numpy as np

FLT_EPSILON = 1.1920929e-07
DBL_EPSILON = 2.220446049250313e-16


class CvRNG:
    """cv::RNG multiply-with-carry generator, seeded with (uint64)-1 as RANSACPointSetRegistrator does."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def _collinear(p):
    i = len(p) - 1
    for j in range(i):
        dx1 = float(p[j][0]) - float(p[i][0]); dy1 = float(p[j][1]) - float(p[i][1])
        for k in range(j):
            dx2 = float(p[k][0]) - float(p[i][0]); dy2 = float(p[k][1]) - float(p[i][1])
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3(p, t):
    a = np.array([[p[t[0]][0], p[t[0]][1], 1.0], [p[t[1]][0], p[t[1]][1], 1.0], [p[t[2]][0], p[t[2]][1], 1.0]], dtype=np.float64)
    return (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
            + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))


def check_subset(s1, s2):
    if _collinear(s1) or _collinear(s2):
        return False
    neg = 0
    for t in ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)):
        neg += (_det3(s1, t) * _det3(s2, t)) < 0
    return neg in (0, 4)


def get_subset(rng, src, dst, max_attempts=10000):
    n = len(src)
    for _ in range(max_attempts):
        idx = []
        for i in range(4):
            v = rng.uniform(0, n)
            while v in idx:
                v = rng.uniform(0, n)
            idx.append(v)
        if check_subset(src[idx], dst[idx]):
            return idx
    return None


def run_kernel(M, m):
    """HomographyEstimatorCallback::runKernel: normalised DLT, smallest eigenvector of LtL.  M -> m.  Returns 3x3 or None."""
    M = M.astype(np.float64); m = m.astype(np.float64)
    n = len(M)
    cM = M.sum(0) / n; cm = m.sum(0) / n
    sM = np.abs(M - cM).sum(0); sm = np.abs(m - cm).sum(0)
    if (np.abs(np.concatenate([sM, sm])) < DBL_EPSILON).any():
        return None
    sM = n / sM; sm = n / sm
    invHnorm = np.array([[1.0 / sm[0], 0, cm[0]], [0, 1.0 / sm[1], cm[1]], [0, 0, 1]])
    Hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
    x = (m[:, 0] - cm[0]) * sm[0]; y = (m[:, 1] - cm[1]) * sm[1]
    X = (M[:, 0] - cM[0]) * sM[0]; Y = (M[:, 1] - cM[1]) * sM[1]
    one = np.ones(n); zero = np.zeros(n)
    Lx = np.stack([X, Y, one, zero, zero, zero, -x * X, -x * Y, -x], 1)
    Ly = np.stack([zero, zero, zero, X, Y, one, -y * X, -y * Y, -y], 1)
    LtL = Lx.T @ Lx + Ly.T @ Ly
    w, V = np.linalg.eigh(LtL)
    H0 = V[:, 0].reshape(3, 3)
    H = invHnorm @ H0 @ Hnorm2
    return H / H[2, 2]


def reproj_err_f32(H, M, m):
    Hf = H.astype(np.float32).ravel()
    Mx, My = M[:, 0].astype(np.float32), M[:, 1].astype(np.float32)
    ww = np.float32(1.0) / (Hf[6] * Mx + Hf[7] * My + np.float32(1.0))
    dx = (Hf[0] * Mx + Hf[1] * My + Hf[2]) * ww - m[:, 0].astype(np.float32)
    dy = (Hf[3] * Mx + Hf[4] * My + Hf[5]) * ww - m[:, 1].astype(np.float32)
    return dx * dx + dy * dy


def update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0); ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, 2.2250738585072014e-308)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < 2.2250738585072014e-308:
        return 0
    num = math.log(num); denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def cv_jacobi(A):
    """cv::eigen of a symmetric matrix (OpenCV core/src/lapack.cpp, JacobiImpl_): classical Jacobi -- the pivot is the largest off-diagonal
    element, found through per-row / per-column maxima that are refreshed for the two rotated indices only; rotation from
    hypot(p, y) with y = (w_l - w_k) / 2; at most 30 n^2 rotations, stop at |pivot| <= DBL_EPSILON; eigenvalues sorted descending by
    selection sort, eigenvectors in ROWS.  Matches cv2.eigen to an ulp (tests/test_oracle_features_cpu.py) and -- what matters here --
    keeps the RELATIVE accuracy of the small eigenvalues, which LAPACK's tridiagonal QR (np.linalg.eigh) loses at condition 1e15."""
    A = np.array(A, dtype=np.float64)
    n = len(A)
    V = np.eye(n)
    W = np.diag(A).copy()
    ind_r = [0] * n
    ind_c = [0] * n

    def row_max(k):
        m = k + 1; mv = abs(A[k, m])
        for i in range(k + 2, n):
            v = abs(A[k, i])
            if mv < v:
                mv = v; m = i
        return m

    def col_max(k):
        m = 0; mv = abs(A[0, k])
        for i in range(1, k):
            v = abs(A[i, k])
            if mv < v:
                mv = v; m = i
        return m

    for k in range(n):
        if k < n - 1:
            ind_r[k] = row_max(k)
        if k > 0:
            ind_c[k] = col_max(k)
    for _ in range(n * n * 30 if n > 1 else 0):
        k = 0; mv = abs(A[0, ind_r[0]])
        for i in range(1, n - 1):
            v = abs(A[i, ind_r[i]])
            if mv < v:
                mv = v; k = i
        l = ind_r[k]
        for i in range(1, n):
            v = abs(A[ind_c[i], i])
            if mv < v:
                mv = v; k = ind_c[i]; l = i
        p = A[k, l]
        if abs(p) <= DBL_EPSILON:
            break
        y = (W[l] - W[k]) * 0.5
        t = abs(y) + math.hypot(p, y)
        s = math.hypot(p, t)
        c = t / s
        s = p / s; t = (p / t) * p
        if y < 0:
            s = -s; t = -t
        A[k, l] = 0.0
        W[k] -= t; W[l] += t
        for (i0, j0, i1, j1) in ([(i, k, i, l) for i in range(0, k)] + [(k, i, i, l) for i in range(k + 1, l)]
                                 + [(k, i, l, i) for i in range(l + 1, n)]):
            a0 = A[i0, j0]; b0 = A[i1, j1]
            A[i0, j0] = a0 * c - b0 * s; A[i1, j1] = a0 * s + b0 * c
        vk = V[k].copy(); vl = V[l].copy()
        V[k] = vk * c - vl * s; V[l] = vk * s + vl * c
        for idx in (k, l):
            if idx < n - 1:
                ind_r[idx] = row_max(idx)
            if idx > 0:
                ind_c[idx] = col_max(idx)
    for k in range(n - 1):
        m = k
        for i in range(k + 1, n):
            if W[m] < W[i]:
                m = i
        if k != m:
            W[[m, k]] = W[[k, m]]; V[[m, k]] = V[[k, m]]
    return W, V


def _eig_pinv(A):
    """What cv::solve / cv::invert do with DECOMP_EIG: symmetric eigen-decomposition (cv_jacobi), then SVBkSb's back substitution,
    which DROPS every eigenvalue with |w_i| <= 2 * DBL_EPSILON * sum(w) -- a truncated pseudo-inverse.  Returns (w, V, keep), eigenvectors
    in the columns of V."""
    w, Vr = cv_jacobi(A)
    keep = np.abs(w) > 2.0 * DBL_EPSILON * w.sum()
    return w, Vr.T, keep


def _solve_eig(A, b):
    w, V, keep = _eig_pinv(A)
    return V[:, keep] @ ((V[:, keep].T @ b) / w[keep])


def _invert_eig_diag(A):
    w, V, keep = _eig_pinv(A)
    return (V[:, keep] ** 2) @ (1.0 / w[keep])


def lm_refine(H, M, m, max_iters=10):
    """LMSolver (calib3d levmarq.cpp: the lambda / lc schedule) with cv2 4.13's HomographyRefineCallback, which refines ALL NINE
    elements of H (fundam.cpp asserts `J.cols == 9`; the classic callback fixed h33 = 1 and had 8 columns) and scales by 1 / h33
    afterwards.  J^T J is therefore singular along the scale gauge h -> (1 + e) h, and what keeps the iteration defined is that
    `solve(Ap, v, d, DECOMP_EIG)` / `invert(A, Ap, DECOMP_EIG)` are truncated pseudo-inverses (_eig_pinv).  On well-conditioned
    consensus sets this converges to the same homography as the 8-parameter form; on ill-conditioned ones (inliers in a corner of the
    frame) the two walk different paths in ten iterations -- pinned by tests/golden/ransac_illcond.npz (frame 359 of clip 01)."""
    M = M.astype(np.float64); m = m.astype(np.float64)

    def compute(h, want_j):
        ww = 1.0 / (h[6] * M[:, 0] + h[7] * M[:, 1] + h[8])
        xi = (h[0] * M[:, 0] + h[1] * M[:, 1] + h[2]) * ww
        yi = (h[3] * M[:, 0] + h[4] * M[:, 1] + h[5]) * ww
        r = np.empty(2 * len(M)); r[0::2] = xi - m[:, 0]; r[1::2] = yi - m[:, 1]
        if not want_j:
            return r, None
        J = np.zeros((2 * len(M), 9))
        J[0::2, 0] = M[:, 0] * ww; J[0::2, 1] = M[:, 1] * ww; J[0::2, 2] = ww
        J[0::2, 6] = -M[:, 0] * ww * xi; J[0::2, 7] = -M[:, 1] * ww * xi; J[0::2, 8] = -ww * xi
        J[1::2, 3] = M[:, 0] * ww; J[1::2, 4] = M[:, 1] * ww; J[1::2, 5] = ww
        J[1::2, 6] = -M[:, 0] * ww * yi; J[1::2, 7] = -M[:, 1] * ww * yi; J[1::2, 8] = -ww * yi
        return r, J

    x = H.ravel().copy()
    r, J = compute(x, True)
    S = float(r @ r)
    A = J.T @ J; v = J.T @ r
    D = np.diag(A).copy()
    Rlo, Rhi = 0.25, 0.75
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A + np.diag(lam * D)
        d = _solve_eig(Ap, v)
        xd = x - d
        rd, _ = compute(xd, False)
        Sd = float(rd @ rd)
        temp = -(A @ d) + 2.0 * v
        dS = float(d @ temp)
        R = (S - Sd) / (dS if abs(dS) > DBL_EPSILON else 1.0)
        if R > Rhi:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < Rlo:
            t = float(d @ v)
            nu = (Sd - S) / (t if abs(t) > DBL_EPSILON else 1.0) + 2.0
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0.0:
                maxval = max(DBL_EPSILON, float(np.abs(_invert_eig_diag(A)).max()))
                lam = lc = 1.0 / maxval
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x = xd
            r, J = compute(x, True)
            A = J.T @ J; v = J.T @ r
        it += 1
        if not (it < max_iters and np.abs(d).max() >= FLT_EPSILON and np.abs(r).max() >= FLT_EPSILON):
            break
    return (x / x[8]).reshape(3, 3)


def find_homography_ransac(src, dst, thresh=2.0, max_iters=2000, confidence=0.995, return_trace=False):
    """src, dst: (n,2) float32 (cur -> prev).  Returns H (3x3 float64) or None; with return_trace also the RANSAC
    trace (accepted subsets in order, inlier counts, iterations run, inlier mask of the winning hypothesis)."""
    src = np.asarray(src, dtype=np.float32).reshape(-1, 2); dst = np.asarray(dst, dtype=np.float32).reshape(-1, 2)
    n = len(src)
    trace = {"subsets": [], "good": [], "iters": 0}
    if n < 4:
        return (None, trace) if return_trace else None
    if n == 4:
        H = run_kernel(src, dst)
        return (H, trace) if return_trace else H
    rng = CvRNG()
    niters = max(max_iters, 1)
    best_good, best_H, best_mask = 0, None, None
    it = 0
    t2 = thresh * thresh
    while it < niters:
        idx = get_subset(rng, src, dst)
        if idx is None:
            if it == 0:
                return (None, trace) if return_trace else None
            break
        Hs = run_kernel(src[idx], dst[idx])
        trace["subsets"].append(idx)
        if Hs is not None:
            mask = reproj_err_f32(Hs, src, dst) <= np.float32(t2)
            good = int(mask.sum())
            trace["good"].append(good)
            if good > max(best_good, 3):
                best_good, best_H, best_mask = good, Hs, mask
                niters = update_num_iters(confidence, (n - good) / n, 4, niters)
        else:
            trace["good"].append(-1)
        it += 1
    trace["iters"] = it
    if best_good <= 0:
        return (None, trace) if return_trace else None
    trace["mask"] = best_mask
    s1, d1 = src[best_mask], dst[best_mask]
    H = run_kernel(s1, d1)
    if H is None:
        H = best_H
    H = lm_refine(H, s1, d1)
    return (H, trace) if return_trace else H
