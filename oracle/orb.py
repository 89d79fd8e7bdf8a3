"""NumPy restatement of cv2.ORB_create(700).detectAndCompute (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference call sites: main.py:36 (`cv2.ORB_create(700)`), :112, :718.  The algorithm is OpenCV 4.x `ORB_Impl`
(features2d/src/orb.cpp) + FAST-9/16 (fast.cpp) + INTER_LINEAR_EXACT resize, restated from the published algorithm as
pinned by SURVEY.md A.2-A.4 and checked against live cv2 4.13 in tests/test_oracle_orb_cpu.py.
Output order differs from cv2 (whose order is whatever libstdc++'s nth_element leaves): here keypoints are
level-major and row-major inside a level; compare as sets (see `canon`).
"""
from __future__ import annotations

import numpy as np

from ._orb_pattern import ORB_PATTERN

N_LEVELS = 8
SCALE_FACTOR = np.float32(1.2)
EDGE = 31
PATCH = 31
HALF_PATCH = 15
FAST_THR = 20
UMAX = np.array([15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3], dtype=np.int64)
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
        (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def cv_round(x):
    return np.rint(x).astype(np.int64)


def level_scales():
    return np.array([np.float32(np.power(np.float64(SCALE_FACTOR), float(l))) for l in range(N_LEVELS)], dtype=np.float32)


def level_sizes(w, h):
    out = []
    for s in level_scales():
        inv = np.float32(1.0) / s
        out.append((int(cv_round(np.float32(w) * inv)), int(cv_round(np.float32(h) * inv))))
    return out


def level_quotas(nfeatures=700):
    factor = np.float32(1.0 / np.float64(SCALE_FACTOR))
    nd = np.float32(nfeatures) * (np.float32(1) - factor) / (np.float32(1) - np.float32(np.power(np.float64(factor), float(N_LEVELS))))
    q, tot = [], 0
    for _ in range(N_LEVELS - 1):
        q.append(int(cv_round(nd)))
        tot += q[-1]
        nd = np.float32(nd * factor)
    q.append(max(nfeatures - tot, 0))
    return q


def resize_linear_exact(src, dw, dh):
    """cv2.resize(src, (dw,dh), interpolation=INTER_LINEAR_EXACT) for uint8 single channel (8.8 fixed-point coefficients)."""
    sh, sw = src.shape

    def coeffs(dn, sn):
        scale = np.float64(sn) / np.float64(dn)
        f = (np.arange(dn, dtype=np.float64) + 0.5) * scale - 0.5
        s = np.floor(f).astype(np.int64)
        a = f - s
        lo = s < 0
        s = np.where(lo, 0, s); a = np.where(lo, 0.0, a)
        hi = s >= sn - 1
        s = np.where(hi, sn - 1, s); a = np.where(hi, 0.0, a)
        a8 = cv_round(a * 256.0)
        return s, np.minimum(s + 1, sn - 1), a8
    sx, sx1, ax = coeffs(dw, sw)
    sy, sy1, ay = coeffs(dh, sh)
    s = src.astype(np.int64)
    hrow = s[:, sx] * (256 - ax)[None, :] + s[:, sx1] * ax[None, :]
    out = (hrow[sy, :] * (256 - ay)[:, None] + hrow[sy1, :] * ay[:, None] + 32768) >> 16
    return out.astype(np.uint8)


def build_pyramid(gray):
    h, w = gray.shape
    levels = [gray]
    for (lw, lh) in level_sizes(w, h)[1:]:
        levels.append(resize_linear_exact(levels[-1], lw, lh))
    return levels


def fast_score_map(img, thr=FAST_THR):
    """score(x,y) for every pixel with a 3 px margin (0 elsewhere / for non-corners): OpenCV cornerScore<16>."""
    h, w = img.shape
    I = img.astype(np.int16)
    c = I[3:h - 3, 3:w - 3]
    d = np.stack([c - I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in RING])          # v - ring
    d2 = np.concatenate([d, d[:8]], axis=0)
    amax = np.full(c.shape, -32768, np.int16)
    bmin = np.full(c.shape, 32767, np.int16)
    for k in range(16):
        arc = d2[k:k + 9]
        amax = np.maximum(amax, arc.min(axis=0))
        bmin = np.minimum(bmin, arc.max(axis=0))
    m = np.maximum(amax, -bmin).astype(np.int32)
    score = np.where(m > thr, m - 1, 0)
    out = np.zeros((h, w), np.int32)
    out[3:h - 3, 3:w - 3] = score
    return out


def fast_nms(score):
    """keep (x,y) iff score > all 8 neighbours (non-corners count 0); row-major order."""
    h, w = score.shape
    p = np.pad(score, 1)
    c = p[1:-1, 1:-1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= c > p[1 + dy:h + 1 + dy, 1 + dx:w + 1 + dx]
    ys, xs = np.nonzero(keep)
    return xs, ys, score[ys, xs]


def retain_best(resp, n):
    """KeyPointsFilter::retainBest: keep everything >= the n-th largest response (ties kept). Boolean mask."""
    if n <= 0:
        return np.zeros(len(resp), bool)
    if len(resp) <= n:
        return np.ones(len(resp), bool)
    thr = np.sort(resp)[::-1][n - 1]
    return resp >= thr


def harris_responses(img, xs, ys):
    I = img.astype(np.int64)
    a = np.zeros(len(xs), np.int64); b = np.zeros(len(xs), np.int64); c = np.zeros(len(xs), np.int64)
    for dy in range(-3, 4):
        for dx in range(-3, 4):
            y, x = ys + dy, xs + dx
            Ix = (I[y, x + 1] - I[y, x - 1]) * 2 + (I[y - 1, x + 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y + 1, x - 1])
            Iy = (I[y + 1, x] - I[y - 1, x]) * 2 + (I[y + 1, x - 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y - 1, x + 1])
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy
    fa, fb, fc = a.astype(np.float32), b.astype(np.float32), c.astype(np.float32)
    scale = np.float32(1.0) / (np.float32(4 * 7) * np.float32(255.0))
    s4 = scale * scale * scale * scale
    return ((fa * fb - fc * fc - np.float32(0.04) * (fa + fb) * (fa + fb)) * s4).astype(np.float32)


def fast_atan2(y, x):
    """cv::fastAtan2 scalar path: float32 Horner, no FMA."""
    y = y.astype(np.float32); x = x.astype(np.float32)
    k = np.float32(180.0 / np.pi)
    p1 = np.float32(0.9997878412794807) * k
    p3 = np.float32(-0.3258083974640975) * k
    p5 = np.float32(0.1555786518463281) * k
    p7 = np.float32(-0.04432655554792128) * k
    ax, ay = np.abs(x), np.abs(y)
    eps = np.float32(2.220446049250313e-16)
    swap = ay > ax
    c = np.where(swap, ax / (ay + eps), ay / (ax + eps)).astype(np.float32)
    c2 = c * c
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(swap, np.float32(90.0) - a, a)
    a = np.where(x < 0, np.float32(180.0) - a, a)
    a = np.where(y < 0, np.float32(360.0) - a, a)
    return a.astype(np.float32)


def ic_angles(img, xs, ys):
    I = img.astype(np.int64)
    m10 = np.zeros(len(xs), np.int64); m01 = np.zeros(len(xs), np.int64)
    for u in range(-HALF_PATCH, HALF_PATCH + 1):
        m10 += u * I[ys, xs + u]
    for v in range(1, HALF_PATCH + 1):
        vs = np.zeros(len(xs), np.int64)
        d = int(UMAX[v])
        for u in range(-d, d + 1):
            p, m = I[ys + v, xs + u], I[ys - v, xs + u]
            vs += p - m
            m10 += u * (p + m)
        m01 += v * vs
    return fast_atan2(m01.astype(np.float32), m10.astype(np.float32))


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def blur7(img):
    """ORB's GaussianBlur(7x7, sigma 2, REFLECT_101) of a level.  Because ORB blurs a sub-matrix of its pyramid image,
    OpenCV skips the u8 fixed-point Gaussian and runs the generic float sepFilter2D, whose AVX2 kernels accumulate
    rows as s = k0*x0; s = fma(k_t, x_t, s) (left to right) and columns in the symmetric form
    s = k3*x3; s = fma(k_{3+t}, x_{3+t} + x_{3-t}, s); the float result is rounded (half-even) to uint8.
    (probed: identical to cv2.sepFilter2D(u8, float kernel) on every pixel, and to cv2.ORB's descriptors.)"""
    x = np.arange(7, dtype=np.float64) - 3
    k = np.exp(-(x * x) / 8.0); k = (k / k.sum()).astype(np.float32)
    h, w = img.shape
    f = np.pad(img.astype(np.float32), 3, mode="reflect")
    row = (k[0] * f[:, 0:w]).astype(np.float32)
    for t in range(1, 7):
        row = _fma32(np.full_like(row, k[t]), f[:, t:t + w], row)
    out = (k[3] * row[3:3 + h, :]).astype(np.float32)
    for t in range(1, 4):
        out = _fma32(np.full_like(out, k[3 + t]), row[3 + t:3 + t + h, :] + row[3 - t:3 - t + h, :], out)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def describe(blurred, xs, ys, angles_deg):
    """rBRIEF-256 at level coordinates (xs,ys) with keypoint angles (degrees)."""
    th = angles_deg.astype(np.float32) * np.float32(np.pi / 180.0)
    a = np.cos(th.astype(np.float64)).astype(np.float32)[:, None]
    b = np.sin(th.astype(np.float64)).astype(np.float32)[:, None]
    px = ORB_PATTERN[:, 0].astype(np.float32)[None, :]
    py = ORB_PATTERN[:, 1].astype(np.float32)[None, :]
    ix = cv_round(px * a - py * b)
    iy = cv_round(px * b + py * a)
    v = blurred[ys[:, None] + iy, xs[:, None] + ix].astype(np.int32)          # (n, 512)
    bits = (v[:, 0::2] < v[:, 1::2]).astype(np.uint8).reshape(len(xs), 32, 8)
    return (bits << np.arange(8, dtype=np.uint8)[None, None, :]).sum(axis=2).astype(np.uint8)


def detect_and_compute(gray, nfeatures=700, blur_fn=blur7):
    """Returns (kp, des): kp float64 (n,7) = x, y, size, angle, response, octave, (level-x | level-y packed for debugging);
    des uint8 (n,32)."""
    levels = build_pyramid(gray)
    scales = level_scales()
    quotas = level_quotas(nfeatures)
    rows, descs = [], []
    for l, img in enumerate(levels):
        h, w = img.shape
        xs, ys, sc = fast_nms(fast_score_map(img))
        inb = (xs >= EDGE) & (xs < w - EDGE) & (ys >= EDGE) & (ys < h - EDGE)
        xs, ys, sc = xs[inb], ys[inb], sc[inb]
        k1 = retain_best(sc.astype(np.float32), 2 * quotas[l])
        xs, ys = xs[k1], ys[k1]
        if len(xs) == 0:
            continue
        resp = harris_responses(img, xs, ys)
        k2 = retain_best(resp, quotas[l])
        xs, ys, resp = xs[k2], ys[k2], resp[k2]
        ang = ic_angles(img, xs, ys)
        s = scales[l]
        ptx = xs.astype(np.float32) * s
        pty = ys.astype(np.float32) * s
        inv = np.float32(1.0) / s
        cx = cv_round(ptx * inv); cy = cv_round(pty * inv)                    # what computeOrbDescriptors re-derives
        descs.append(describe(blur_fn(img), cx, cy, ang))
        n = len(xs)
        rows.append(np.stack([ptx.astype(np.float64), pty.astype(np.float64), np.full(n, np.float64(np.float32(PATCH) * s)),
                              ang.astype(np.float64), resp.astype(np.float64), np.full(n, float(l))], axis=1))
    if not rows:
        return np.zeros((0, 6)), np.zeros((0, 32), np.uint8)
    return np.concatenate(rows), np.concatenate(descs)


def canon(kp, des):
    """canonical order for set comparison: sort by (octave, y, x)."""
    o = np.lexsort((kp[:, 0], kp[:, 1], kp[:, 5]))
    return kp[o], des[o]


def cv_detect_and_compute(gray, nfeatures=700):
    """live cv2 (the reference's actual call), converted to the same array form."""
    import cv2
    kps, des = cv2.ORB_create(nfeatures).detectAndCompute(gray, None)
    kp = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], dtype=np.float64).reshape(-1, 6)
    return kp, (des if des is not None else np.zeros((0, 32), np.uint8))
