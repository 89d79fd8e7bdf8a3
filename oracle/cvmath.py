"""NumPy restatements of the OpenCV primitives on the stitching path (TEST INFRASTRUCTURE, see oracle/__init__.py).

OpenCV itself is a third-party dependency of the reference (requirements.txt:1-2, unpinned; 4.13.0 in this
image) and is not under /root/reference; each function restates the published OpenCV 4.x algorithm for the call
the reference makes, and is pinned against live cv2 in tests/test_oracle_cpu.py.
Call sites in the reference are cited per function (file main.py).
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------------------------------------
# cv2.cvtColor(BGR2GRAY)   main.py:111,717     (15-bit fixed point, exact)
# ------------------------------------------------------------------------------------------------
def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# cv2.warpPerspective(frame, H, (Wc,Hc), INTER_LINEAR)   main.py:871   
# ------------------------------------------------------------------------------------------------
def invert3x3(H: np.ndarray) -> np.ndarray:
    """Closed-form cofactor inverse in double, the association OpenCV's cv::invert uses for 3x3."""
    h = np.asarray(H, dtype=np.float64)
    a = h.ravel()
    det = (a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6])
           + a[2] * (a[3] * a[7] - a[4] * a[6]))
    if det == 0.0:
        return np.zeros((3, 3))
    d = 1.0 / det
    t = np.empty(9)
    t[0] = (a[4] * a[8] - a[5] * a[7]) * d
    t[1] = (a[2] * a[7] - a[1] * a[8]) * d
    t[2] = (a[1] * a[5] - a[2] * a[4]) * d
    t[3] = (a[5] * a[6] - a[3] * a[8]) * d
    t[4] = (a[0] * a[8] - a[2] * a[6]) * d
    t[5] = (a[2] * a[3] - a[0] * a[5]) * d
    t[6] = (a[3] * a[7] - a[4] * a[6]) * d
    t[7] = (a[1] * a[6] - a[0] * a[7]) * d
    t[8] = (a[0] * a[4] - a[1] * a[3]) * d
    return t.reshape(3, 3)


def warp_perspective(src: np.ndarray, H: np.ndarray, dsize) -> np.ndarray:
    """INTER_LINEAR, BORDER_CONSTANT(0), INTER_BITS=5 fixed point, evaluated in OpenCV's 64-column blocks."""
    Wc, Hc = dsize
    M = invert3x3(H).ravel()
    sh, sw = src.shape[:2]
    x = np.arange(Wc, dtype=np.int64)[None, :]
    y = np.arange(Hc, dtype=np.float64)[:, None]
    bx = (x // 64) * 64
    x1 = (x - bx).astype(np.float64)
    bxf = bx.astype(np.float64)
    X0 = (M[0] * bxf + M[1] * y) + M[2]
    Y0 = (M[3] * bxf + M[4] * y) + M[5]
    W0 = (M[6] * bxf + M[7] * y) + M[8]
    W = W0 + M[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        Wi = np.where(W != 0, 32.0 / W, 0.0)
    fX = np.clip((X0 + M[0] * x1) * Wi, -2147483648.0, 2147483647.0)
    fY = np.clip((Y0 + M[3] * x1) * Wi, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    # OpenCV stores the integer source coordinates as saturated int16
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    ax = (X & 31)
    ay = (Y & 31)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return v * ok[..., None]

    w00 = ((32 - ax) * (32 - ay))[..., None]
    w01 = (ax * (32 - ay))[..., None]
    w10 = ((32 - ax) * ay)[..., None]
    w11 = (ax * ay)[..., None]
    out = (tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11 + 512) >> 10
    return out.astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# cv2.distanceTransform(mask, DIST_L2, 3)   main.py:888-889   (IPP off: integer chamfer)
# ------------------------------------------------------------------------------------------------
CHAMFER_A = 62587        # cvRound(0.955f  * 65536)
CHAMFER_B = 89738        # cvRound(1.3693f * 65536)
CHAMFER_INIT = (2 ** 31 - 1) >> 2


def chamfer_dt_int(mask: np.ndarray) -> np.ndarray:
    """Exact integer chamfer distances (16.16 fixed point) = result of OpenCV's two raster passes, computed with the
    same forward/backward recurrences (rows sequential, the in-row term as a prefix-min scan).  int64, capped at
    CHAMFER_INIT like OpenCV's DIST_MAX."""
    a, b = CHAMFER_A, CHAMFER_B
    Hh, Ww = mask.shape
    INF = np.int64(CHAMFER_INIT)
    big = np.int64(1) << 40
    xs = np.arange(Ww, dtype=np.int64) * a
    nz = mask != 0
    d = np.empty((Hh, Ww), dtype=np.int64)
    prev = np.full(Ww + 2, INF, dtype=np.int64)
    for yy in range(Hh):                                   # forward pass
        up = np.minimum(np.minimum(prev[:-2] + b, prev[2:] + b), prev[1:-1] + a)
        t = np.where(nz[yy], np.minimum(up, INF + a), 0)   # left border neighbour is INIT
        # in-row: d[x] = min_k<=x t[k] + a*(x-k), but a zero pixel resets to 0 (already in t)
        d[yy] = np.minimum.accumulate(t - xs) + xs
        d[yy] = np.minimum(d[yy], big)
        prev[1:-1] = d[yy]
    prev[:] = INF
    for yy in range(Hh - 1, -1, -1):                       # backward pass
        dn = np.minimum(np.minimum(prev[:-2] + b, prev[2:] + b), prev[1:-1] + a)
        t = np.minimum(d[yy], dn)
        r = (np.minimum.accumulate((t + xs)[::-1])[::-1]) - xs
        d[yy] = np.minimum(t, r)
        prev[1:-1] = d[yy]
    return np.minimum(d, INF)


def chamfer_dt(mask: np.ndarray) -> np.ndarray:
    """float32 output exactly as OpenCV: (float)t * (1.f/65536)."""
    return chamfer_dt_int(mask).astype(np.float32) * np.float32(1.0 / 65536.0)


def chamfer_dt_closed_form(mask: np.ndarray) -> np.ndarray:
    """Closed form D(p) = min_q a*max(|dx|,|dy|) + (b-a)*min(|dx|,|dy|) over zero pixels q via the row/column
    decomposition.  O(H^2 W): small masks only.  Used to pin the decomposition the CUDA path uses."""
    a, b = CHAMFER_A, CHAMFER_B
    Hh, Ww = mask.shape
    BIG = 1 << 20
    g = np.full((Hh, Ww), BIG, dtype=np.int64)
    for yy in range(Hh):
        z = np.flatnonzero(mask[yy] == 0)
        if z.size:
            xs = np.arange(Ww)
            idx = np.searchsorted(z, xs)
            left = np.where(idx > 0, xs - z[np.clip(idx - 1, 0, z.size - 1)], BIG)
            right = np.where(idx < z.size, z[np.clip(idx, 0, z.size - 1)] - xs, BIG)
            g[yy] = np.minimum(left, right)
    out = np.full((Hh, Ww), CHAMFER_INIT, dtype=np.int64)
    ys = np.arange(Hh)
    for yy in range(Hh):
        v = np.abs(ys - yy)[:, None]
        cost = a * np.maximum(g, v) + (b - a) * np.minimum(g, v)
        cost = np.where(g >= BIG, CHAMFER_INIT, cost)
        out[yy] = np.minimum(cost.min(axis=0), CHAMFER_INIT)
    return out


# ------------------------------------------------------------------------------------------------
# cv2.GaussianBlur(w, (31,31), 0) on float32   main.py:897-898   (sigma = 0.3*((31-1)*0.5-1)+0.8 = 5.0)
# ------------------------------------------------------------------------------------------------
def gaussian_kernel_f32(ksize: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksize, sigma, CV_32F): exp in double, normalised in double, cast to float."""
    r = (ksize - 1) * 0.5
    x = np.arange(ksize, dtype=np.float64) - r
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    k = k / k.sum()
    return k.astype(np.float32)


def reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def _fma32(a, b, c):
    """float32 fused multiply-add emulated through float64 (the product of two float32 is exact in float64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def blur31(img: np.ndarray, kernel: np.ndarray | None = None) -> np.ndarray:
    """Separable 31-tap float32 filter, BORDER_REFLECT_101, in the accumulation order of OpenCV's AVX2 sepFilter2D
    (probed against cv2 4.13): rows  s = k0*x0; s = fma(k_t, x_t, s) for t = 1..30 (left to right);
    columns s = k_c*x_c; s = fma(k_{c+t}, x_{c+t} + x_{c-t}, s) for t = 1..15 (symmetric form).
    (cv2's scalar tail for the last width%8 columns of the column pass is not contracted; not modelled: <= 1 ulp.)"""
    k = gaussian_kernel_f32(31, 5.0) if kernel is None else kernel
    Hh, Ww = img.shape
    R = len(k) // 2
    xi = reflect101(np.arange(-R, Ww + R), Ww)
    yi = reflect101(np.arange(-R, Hh + R), Hh)
    src = img.astype(np.float32)
    pad = src[:, xi]
    row = (k[0] * pad[:, 0:Ww]).astype(np.float32)
    for t in range(1, len(k)):
        row = _fma32(np.full_like(row, k[t]), pad[:, t:t + Ww], row)
    padv = row[yi, :]
    out = (k[R] * padv[R:R + Hh]).astype(np.float32)
    for t in range(1, R + 1):
        out = _fma32(np.full_like(out, k[R + t]), padv[R + t:R + t + Hh] + padv[R - t:R - t + Hh], out)
    return out


# ------------------------------------------------------------------------------------------------
# blend step of VideMosaic.warp   main.py:878-927   (SURVEY A.10), all restated primitives
# ------------------------------------------------------------------------------------------------
def blend_step(canvas_u8: np.ndarray, warped_u8: np.ndarray) -> np.ndarray:
    mask_new = np.any(warped_u8 > 0, axis=2)
    mask_old = np.any(canvas_u8 > 0, axis=2)
    overlap = mask_new & mask_old
    if overlap.any():
        dn = chamfer_dt(mask_new.astype(np.uint8))
        do = chamfer_dt(mask_old.astype(np.uint8))
        s = (dn + do) + np.float32(1e-6)
        wn = blur31(dn / s)
        wo = blur31(do / s)
        blended = canvas_u8.astype(np.float32) * wo[..., None] + warped_u8.astype(np.float32) * wn[..., None]
        out = np.where(overlap[..., None], blended.astype(np.uint8), canvas_u8)
        out = np.where((mask_new & ~overlap)[..., None], warped_u8, out)
        return out
    out = canvas_u8.copy()
    sel = warped_u8 > 0
    out[sel] = warped_u8[sel]
    return out
