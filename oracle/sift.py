"""Restatement of the front half of cv2.SIFT_create(700).detectAndCompute (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference call sites: main.py:33 (`cv2.SIFT_create(700)`), :112, :718.  OpenCV 4.x features2d/sift: nOctaveLayers 3,
contrastThreshold 0.04, edgeThreshold 10, sigma 1.6, first octave -1.  Restated here: the initial image,
the Gaussian / DoG pyramids (as the same sequence of cv2 primitive calls SIFT makes internally) and the scale-space
extrema + sub-pixel refinement in NumPy.  Orientation assignment and the 128-d descriptor are checked against live cv2
SIFT output directly (tests/test_sift_gpu.py) with the tolerances stated there; `match_keypoints` does the set matching."""
from __future__ import annotations

import numpy as np
import cv2

N_LAYERS = 3
SIGMA = 1.6
BORDER = 5


def level_sigmas():
    k = 2.0 ** (1.0 / N_LAYERS)
    out = [SIGMA]
    for i in range(1, N_LAYERS + 3):
        sp = k ** (i - 1) * SIGMA
        st = sp * k
        out.append(float(np.sqrt(st * st - sp * sp)))
    return out


def initial_image(gray):
    """createInitialImage: float32, 2x INTER_LINEAR upsample, blur with sqrt(sigma^2 - 4*0.5^2)."""
    g = gray.astype(np.float32)
    sig_diff = float(np.sqrt(np.float32(max(np.float32(SIGMA * SIGMA) - np.float32(0.5 * 0.5 * 4), 0.01))))
    dbl = cv2.resize(g, (gray.shape[1] * 2, gray.shape[0] * 2), interpolation=cv2.INTER_LINEAR)
    return cv2.GaussianBlur(dbl, (0, 0), sigmaX=sig_diff, sigmaY=sig_diff)


def num_octaves(base):
    return int(np.rint(np.log(float(min(base.shape))) / np.log(2.0) - 2)) + 1


def build_pyramids(gray):
    base = initial_image(gray)
    sig = level_sigmas()
    nO = num_octaves(base)
    gpyr, dpyr = [], []
    for o in range(nO):
        lv = []
        for i in range(N_LAYERS + 3):
            if o == 0 and i == 0:
                lv.append(base)
            elif i == 0:
                src = gpyr[o - 1][N_LAYERS]
                lv.append(cv2.resize(src, (src.shape[1] // 2, src.shape[0] // 2), interpolation=cv2.INTER_NEAREST))
            else:
                lv.append(cv2.GaussianBlur(lv[i - 1], (0, 0), sigmaX=sig[i], sigmaY=sig[i]))
        gpyr.append(lv)
        dpyr.append([lv[i + 1] - lv[i] for i in range(N_LAYERS + 2)])
    return gpyr, dpyr


def _adjust(dog, o, layer, r, c):
    img_scale = np.float32(1.0 / 255); ds = img_scale * np.float32(0.5); sds = img_scale; cds = img_scale * np.float32(0.25)
    h, w = dog[o][0].shape
    xi = xr = xc = np.float32(0)
    for it in range(5):
        img, prv, nxt = dog[o][layer], dog[o][layer - 1], dog[o][layer + 1]
        dD = np.array([(img[r, c + 1] - img[r, c - 1]) * ds, (img[r + 1, c] - img[r - 1, c]) * ds, (nxt[r, c] - prv[r, c]) * ds], np.float32)
        v2 = img[r, c] * np.float32(2)
        dxx = (img[r, c + 1] + img[r, c - 1] - v2) * sds; dyy = (img[r + 1, c] + img[r - 1, c] - v2) * sds
        dss = (nxt[r, c] + prv[r, c] - v2) * sds
        dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cds
        dxs = (nxt[r, c + 1] - nxt[r, c - 1] - prv[r, c + 1] + prv[r, c - 1]) * cds
        dys = (nxt[r + 1, c] - nxt[r - 1, c] - prv[r + 1, c] + prv[r - 1, c]) * cds
        Hm = np.array([[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]], np.float64)
        try:
            X = np.linalg.solve(Hm, dD.astype(np.float64)).astype(np.float32)
        except np.linalg.LinAlgError:
            X = np.zeros(3, np.float32)
        xi, xr, xc = -X[2], -X[1], -X[0]
        if abs(xi) < 0.5 and abs(xr) < 0.5 and abs(xc) < 0.5:
            break
        if max(abs(xi), abs(xr), abs(xc)) > (2 ** 31 - 1) / 3:
            return None
        c += int(np.rint(xc)); r += int(np.rint(xr)); layer += int(np.rint(xi))
        if layer < 1 or layer > N_LAYERS or c < BORDER or c >= w - BORDER or r < BORDER or r >= h - BORDER:
            return None
    else:
        return None
    img, prv, nxt = dog[o][layer], dog[o][layer - 1], dog[o][layer + 1]
    dD = np.array([(img[r, c + 1] - img[r, c - 1]) * ds, (img[r + 1, c] - img[r - 1, c]) * ds, (nxt[r, c] - prv[r, c]) * ds], np.float32)
    t = dD[0] * xc + dD[1] * xr + dD[2] * xi
    contr = img[r, c] * img_scale + t * np.float32(0.5)
    if abs(contr) * N_LAYERS < 0.04:
        return None
    v2 = img[r, c] * np.float32(2)
    dxx = (img[r, c + 1] + img[r, c - 1] - v2) * sds; dyy = (img[r + 1, c] + img[r - 1, c] - v2) * sds
    dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cds
    tr = dxx + dyy; det = dxx * dyy - dxy * dxy
    if det <= 0 or tr * tr * 10 >= 121 * det:
        return None
    size = np.float32(SIGMA) * np.float32(2.0 ** ((layer + xi) / N_LAYERS)) * (1 << o) * 2
    return ((c + xc) * (1 << o), (r + xr) * (1 << o), float(size), float(abs(contr)), o, layer, r, c)


def refined_extrema(dog):
    """all refined scale-space extrema (before orientation assignment): rows (x, y, size, response, octave, layer) in the
    doubled-image coordinate frame; duplicates (same refined cell) removed."""
    out, seen = [], set()
    for o in range(len(dog)):
        h, w = dog[o][0].shape
        if h <= 2 * BORDER or w <= 2 * BORDER:
            continue
        st = np.stack(dog[o])                                  # (5, h, w)
        for layer in range(1, N_LAYERS + 1):
            cur = st[layer, BORDER:h - BORDER, BORDER:w - BORDER]
            nb = []
            for dl in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        if dl == 0 and dy == 0 and dx == 0:
                            continue
                        nb.append(st[layer + dl, BORDER + dy:h - BORDER + dy, BORDER + dx:w - BORDER + dx])
            nb = np.stack(nb)
            ismax = (cur > 0) & (cur >= nb.max(axis=0)); ismin = (cur < 0) & (cur <= nb.min(axis=0))
            ys, xs = np.nonzero((np.abs(cur) > 1.0) & (ismax | ismin))
            for y, x in zip(ys, xs):
                res = _adjust(dog, o, layer, int(y) + BORDER, int(x) + BORDER)
                if res is None:
                    continue
                key = (res[4], res[5], res[6], res[7])
                if key in seen:
                    continue
                seen.add(key)
                out.append(res[:6])
    return np.array(out, dtype=np.float64).reshape(-1, 6)


def cv_detect_and_compute(gray, nfeatures=700):
    kps, des = cv2.SIFT_create(nfeatures).detectAndCompute(gray, None)
    kp = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], dtype=np.float64).reshape(-1, 6)
    return kp, (des if des is not None else np.zeros((0, 128), np.float32))


def match_keypoints(a, b, tol=0.02):
    """greedy one-to-one matching of keypoint rows (x, y, size, angle, ...) by position+size, then angle.  Returns index
    pairs (ia, ib)."""
    pairs, used = [], set()
    for i, r in enumerate(a):
        d = np.abs(b[:, 0] - r[0]) + np.abs(b[:, 1] - r[1]) + np.abs(b[:, 2] - r[2])
        cand = np.nonzero(d < tol)[0]
        best, bj = 1e9, -1
        for j in cand:
            if j in used:
                continue
            da = abs(((b[j, 3] - r[3]) + 180.0) % 360.0 - 180.0)
            if da < best:
                best, bj = da, j
        if bj >= 0:
            used.add(bj)
            pairs.append((i, bj))
    return np.array(pairs, dtype=np.int64).reshape(-1, 2)
