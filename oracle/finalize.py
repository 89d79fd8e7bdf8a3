"""oracle/finalize.py -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the mosaic finalisation that follows the stitching
path: `crop_black_areas` (/root/reference/main.py:980-1003) and `scale_to_screen` (:1006-1038) as
`main()` calls them (:1647-1659: threshold=80, margin=30, screen 1920x1080 off Windows).

Two layers, like the rest of the oracle:
  * crop_black_areas / scale_to_screen: the reference functions restated call for call (cv2 + NumPy), pinned against the
    unmodified reference by tests/golden/finalize.npz (tests/golden/make_golden_finalize.py);
  * crop_rect / resize_linear_u8: the arithmetic itself (third-party OpenCV 4.13.0, absent from /root/reference), restated
    from the published algorithm and checked bit-exact against live cv2 (tests/test_oracle_finalize_cpu.py):
      - BGR2GRAY (3735 B + 19235 G + 9798 R + 16384) >> 15, THRESH_BINARY (> thr), bounding rectangle of the non-zero pixels;
      - cv2.resize(..., INTER_LINEAR) on 8-bit: scale = 1 / (dsize / ssize) in double; per axis
        f = (float)((d + 0.5) * scale - 0.5), s = floor(f), f -= s; HORIZONTALLY s < 0 -> (0, 0), s >= n-1 -> (n-1, 0);
        VERTICALLY only the row indices are clamped, the weights are kept; weights = saturate_cast<short>(w * 2048);
        rows: S = p[s] * a0 + p[s+1] * a1 (int32); columns (the SIMD form every x86 build takes):
        ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        exact 2x2 decimation (ssize == 2 * dsize on both axes) is routed to INTER_AREA: (a + b + c + d + 2) >> 2.
"""
import numpy as np
import cv2


def crop_black_areas(image, threshold=15, margin=5):
    """main.py:980-1003, call for call."""
    if image.dtype != np.uint8:
        image = image.astype(np.uint8)
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    _, thresh = cv2.threshold(gray, threshold, 255, cv2.THRESH_BINARY)
    coords = cv2.findNonZero(thresh)
    if coords is None:
        return image
    x, y, w, h = cv2.boundingRect(coords)
    x = max(0, x + margin)
    y = max(0, y + margin)
    w = min(image.shape[1] - x, w - 2 * margin)
    h = min(image.shape[0] - y, h - 2 * margin)
    return image[y:y + h, x:x + w]


def screen_size(iw, ih, target_w=None, target_h=None):
    """main.py:1011-1036 off Windows (the ctypes.windll lookup raises -> 1920x1080): the size scale_to_screen resizes to."""
    screen_w, screen_h = (1920, 1080) if target_w is None or target_h is None else (target_w, target_h)
    scale = min(screen_w / float(iw), screen_h / float(ih))
    if scale <= 0:
        scale = 1.0
    return max(1, int(iw * scale)), max(1, int(ih * scale))


def scale_to_screen(image, target_w=None, target_h=None):
    """main.py:1006-1038 (non-Windows branch)."""
    ih, iw = image.shape[0], image.shape[1]
    new_w, new_h = screen_size(iw, ih, target_w, target_h)
    return cv2.resize(image, (new_w, new_h), interpolation=cv2.INTER_LINEAR)


# ---------------------------------------------------------------------------------------------------------------
# the arithmetic
# ---------------------------------------------------------------------------------------------------------------
def crop_rect(image, threshold, margin):
    """(x, y, w, h) of crop_black_areas' slice (w or h may be <= 0: an empty slice), or None when nothing is above threshold."""
    b, g, r = (image[..., i].astype(np.int64) for i in range(3))
    gray = (3735 * b + 19235 * g + 9798 * r + 16384) >> 15
    ys, xs = np.nonzero(gray > threshold)
    if len(xs) == 0:
        return None
    x, y, w, h = int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)
    x = max(0, x + margin)
    y = max(0, y + margin)
    w = min(image.shape[1] - x, w - 2 * margin)
    h = min(image.shape[0] - y, h - 2 * margin)
    return x, y, w, h


def _axis(dn, sn, vertical):
    scale = 1.0 / (dn / float(sn))
    idx = np.zeros(dn, np.int64)
    wgt = np.zeros((dn, 2), np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if not vertical:
            if s < 0:
                s, f = 0, np.float32(0.0)
            if s >= sn - 1:
                s, f = sn - 1, np.float32(0.0)
        idx[d] = s
        wgt[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048.0))))
        wgt[d, 1] = int(np.rint(np.float32(f * np.float32(2048.0))))
    return idx, wgt


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 images, bit-exact (cv2 4.13.0, x86 SIMD build)."""
    sh, sw = src.shape[:2]
    s = src.astype(np.int64)
    if sw == 2 * dw and sh == 2 * dh:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xi, xa = _axis(dw, sw, False)
    yi, ya = _axis(dh, sh, True)
    x1 = np.minimum(xi + 1, sw - 1)
    H = s[:, xi] * xa[None, :, 0, None] + s[:, x1] * xa[None, :, 1, None]
    y0 = np.clip(yi, 0, sh - 1)
    y1 = np.clip(yi + 1, 0, sh - 1)
    b0 = ya[:, 0][:, None, None]
    b1 = ya[:, 1][:, None, None]
    out = (((b0 * (H[y0] >> 4)) >> 16) + ((b1 * (H[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def finalize(canvas, threshold=80, margin=30, target_w=None, target_h=None):
    """What main() saves as mosaic.jpg (main.py:1647-1659), restated on the arithmetic layer.  Returns (image, rect)."""
    rect = crop_rect(canvas, threshold, margin)
    x, y, w, h = (0, 0, canvas.shape[1], canvas.shape[0]) if rect is None else rect
    crop = canvas[y:y + h, x:x + w]
    nw, nh = screen_size(crop.shape[1], crop.shape[0], target_w, target_h)
    return resize_linear_u8(crop, nw, nh), (x, y, crop.shape[1], crop.shape[0])
