"""Stage entry points of the C ABI on torch CUDA tensors (plumbing for the parity tests and benches).
Each function cites the reference call it replaces; there is no CPU path."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t: torch.Tensor):
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ingest_bgr(bgr: torch.Tensor):
    """cv2.cvtColor(frame, BGR2GRAY) (main.py:111,717) + BGRX copy.  bgr: (H,W,3) uint8 cuda."""
    lib = _lib.load()
    h, w, _ = bgr.shape
    gray = torch.empty((h, w), dtype=torch.uint8, device=bgr.device)
    bgrx = torch.empty((h, w, 4), dtype=torch.uint8, device=bgr.device)
    _lib.check(lib.bm_ingest_bgr(_ptr(bgr), h, w, _ptr(gray), _ptr(bgrx), _stream()), "bm_ingest_bgr")
    return gray, bgrx


def warp_perspective(src: torch.Tensor, H, dsize):
    """cv2.warpPerspective(src, H, dsize, flags=INTER_LINEAR) (main.py:871).  src (H,W,3) uint8 cuda; dsize=(Wc,Hc)."""
    lib = _lib.load()
    sh, sw, _ = src.shape
    dw, dh = dsize
    dst = torch.empty((dh, dw, 3), dtype=torch.uint8, device=src.device)
    _a, hp = _lib.dbl9(H)
    _lib.check(lib.bm_warp_perspective_bgr(_ptr(src), sh, sw, hp, _ptr(dst), dh, dw, _stream()), "bm_warp_perspective_bgr")
    return dst


def distance_transform(mask: torch.Tensor):
    """cv2.distanceTransform(mask, DIST_L2, 3) (main.py:888-889).  mask (H,W) uint8 cuda -> float32."""
    lib = _lib.load()
    h, w = mask.shape
    out = torch.empty((h, w), dtype=torch.float32, device=mask.device)
    _lib.check(lib.bm_distance_transform(_ptr(mask), h, w, _ptr(out), _stream()), "bm_distance_transform")
    return out


def gaussian_blur31(img: torch.Tensor):
    """cv2.GaussianBlur(img, (31,31), 0) on float32 (main.py:897-898)."""
    lib = _lib.load()
    h, w = img.shape
    out = torch.empty_like(img)
    _lib.check(lib.bm_gaussian_blur31(_ptr(img), h, w, _ptr(out), _stream()), "bm_gaussian_blur31")
    return out


def blend_step(canvas: torch.Tensor, warped: torch.Tensor, win=None):
    """The blend of VideMosaic.warp (main.py:878-927).  canvas, warped: (Hc,Wc,3) uint8 cuda.  Returns (new canvas,
    any_overlap)."""
    lib = _lib.load()
    dh, dw, _ = canvas.shape
    out = canvas.clone()
    flag = C.c_int(0)
    wp = None
    if win is not None:
        wp = (C.c_int * 4)(*[int(v) for v in win])
    _lib.check(lib.bm_blend_step_bgr(_ptr(out), _ptr(warped), dh, dw, wp, C.byref(flag), _stream()), "bm_blend_step_bgr")
    return out, bool(flag.value)


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def orb_detect_and_compute(gray: torch.Tensor, nfeatures: int = 700):
    """cv2.ORB_create(nfeatures).detectAndCompute(gray, None) (main.py:36,112,718).  gray (H,W) uint8 cuda.
    Returns (kp float32 (n,6): x,y,size,angle,response,octave; des uint8 (n,32)) in cv2's own order."""
    lib = _lib.load()
    h, w = gray.shape
    cap = lib.bm_keypoint_capacity()
    kp = np.empty((cap, 6), np.float32); des = np.empty((cap, 32), np.uint8); n = C.c_int(0)
    torch.cuda.synchronize()
    _lib.check(lib.bm_orb_detect_and_compute(_ptr(gray), h, w, nfeatures, _np_ptr(kp), _np_ptr(des), cap, C.byref(n)),
               "bm_orb_detect_and_compute")
    return kp[:n.value].copy(), des[:n.value].copy()


def sift_detect_and_compute(gray: torch.Tensor, nfeatures: int = 700):
    """cv2.SIFT_create(nfeatures).detectAndCompute(gray, None) (main.py:33,112,718). des float32 (n,128)."""
    lib = _lib.load()
    h, w = gray.shape
    cap = lib.bm_keypoint_capacity()
    kp = np.empty((cap, 6), np.float32); des = np.empty((cap, 128), np.float32); n = C.c_int(0)
    torch.cuda.synchronize()
    _lib.check(lib.bm_sift_detect_and_compute(_ptr(gray), h, w, nfeatures, _np_ptr(kp), _np_ptr(des), cap, C.byref(n)),
               "bm_sift_detect_and_compute")
    return kp[:n.value].copy(), des[:n.value].copy()


def cv_retain_best(resp: np.ndarray, n_points: int, as_u8: bool = False) -> np.ndarray:
    """KeyPointsFilter::retainBest (inside detectAndCompute, main.py:112,718): surviving input indices in cv2's output order."""
    lib = _lib.load()
    r = np.ascontiguousarray(resp, np.float32)
    out = np.empty(max(len(r), 1), np.int32); m = C.c_int(0)
    _lib.check(lib.bm_cv_retain_best(_np_ptr(r), len(r), int(n_points), int(as_u8), _np_ptr(out), C.byref(m)), "bm_cv_retain_best")
    return out[:m.value].astype(np.int64)


def jpeg_encode(bgr: np.ndarray, quality: int = 95, device: int = 0) -> bytes:
    """the bytes cv2.imwrite(path, bgr) / cv2.imencode('.jpg', bgr) produce (main.py:1664-1665 writes mosaic.jpg with the default
    quality 95), encoded on the device: baseline JPEG, 4:2:0, Annex K tables, byte for byte cv2 4.13's libjpeg output."""
    lib = _lib.load()
    img = np.ascontiguousarray(bgr)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("jpeg_encode: expected an (h, w, 3) uint8 BGR image")
    h, w = img.shape[:2]
    out = np.empty(lib.bm_jpeg_bound(w, h), np.uint8); n = C.c_size_t(0)
    _lib.check(lib.bm_jpeg_encode(_np_ptr(img), w, h, int(quality), int(device), _np_ptr(out), out.nbytes, C.byref(n)), "bm_jpeg_encode")
    return out[:n.value].tobytes()


def match_hamming_crosscheck(des_q: np.ndarray, des_t: np.ndarray):
    """BFMatcher(NORM_HAMMING, crossCheck=True).match + sorted(key=distance) (main.py:694-698). -> (m,3) float64."""
    lib = _lib.load()
    q = np.ascontiguousarray(des_q, np.uint8); t = np.ascontiguousarray(des_t, np.uint8)
    cap = max(len(q), 1)
    oq = np.empty(cap, np.int32); ot = np.empty(cap, np.int32); od = np.empty(cap, np.float32); m = C.c_int(0)
    _lib.check(lib.bm_match_hamming_crosscheck(_np_ptr(q), len(q), _np_ptr(t), len(t), _np_ptr(oq), _np_ptr(ot), _np_ptr(od),
                                               C.byref(m)), "bm_match_hamming_crosscheck")
    k = m.value
    return np.stack([oq[:k], ot[:k], od[:k]], axis=1).astype(np.float64)


def match_l2_ratio(des_q: np.ndarray, des_t: np.ndarray, ratio: float = 0.7):
    """BFMatcher().knnMatch(k=2) + Lowe ratio + sorted(key=distance) (main.py:687-698). -> (m,3) float64."""
    lib = _lib.load()
    q = np.ascontiguousarray(des_q, np.float32); t = np.ascontiguousarray(des_t, np.float32)
    cap = max(len(q), 1)
    oq = np.empty(cap, np.int32); ot = np.empty(cap, np.int32); od = np.empty(cap, np.float32); m = C.c_int(0)
    _lib.check(lib.bm_match_l2_knn2_ratio(_np_ptr(q), len(q), _np_ptr(t), len(t), float(ratio), _np_ptr(oq), _np_ptr(ot),
                                          _np_ptr(od), C.byref(m)), "bm_match_l2_knn2_ratio")
    k = m.value
    return np.stack([oq[:k], ot[:k], od[:k]], axis=1).astype(np.float64)


def ransac_homography(src: np.ndarray, dst: np.ndarray, thresh: float = 2.0, max_iters: int = 2000, confidence: float = 0.995):
    """cv2.findHomography(src, dst, cv2.RANSAC, thresh) (main.py:856-857).  Returns (H or None, iters, n_inliers)."""
    lib = _lib.load()
    s = np.ascontiguousarray(np.asarray(src, np.float32).reshape(-1, 2)); d = np.ascontiguousarray(np.asarray(dst, np.float32).reshape(-1, 2))
    H = np.zeros(9, np.float64); ok = C.c_int(0); it = C.c_int(0); ni = C.c_int(0)
    _lib.check(lib.bm_ransac_homography(_np_ptr(s), _np_ptr(d), len(s), float(thresh), int(max_iters), float(confidence),
                                        H.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ok), C.byref(it), C.byref(ni)),
               "bm_ransac_homography")
    return (H.reshape(3, 3) if ok.value else None), it.value, ni.value


def orb_debug_level(gray: torch.Tensor, level: int):
    """(level image, FAST score map) of the device ORB pyramid -- parity probe for the INTER_LINEAR_EXACT chain / FAST."""
    lib = _lib.load()
    h, w = gray.shape
    img = np.empty(h * w, np.uint8); sc = np.empty(h * w, np.uint8); lw = C.c_int(0); lh = C.c_int(0)
    torch.cuda.synchronize()
    _lib.check(lib.bm_orb_debug_level(_ptr(gray), h, w, level, _np_ptr(img), _np_ptr(sc), C.byref(lw), C.byref(lh)), "bm_orb_debug_level")
    n = lw.value * lh.value
    return img[:n].reshape(lh.value, lw.value).copy(), sc[:n].reshape(lh.value, lw.value).copy()


def sift_debug_level(gray: torch.Tensor, octave: int, level: int, dog: bool = False):
    """One Gaussian / DoG image of the device SIFT pyramid (float32) and the octave count -- parity probe."""
    lib = _lib.load()
    h, w = gray.shape
    buf = np.empty(4 * h * w, np.float32); lw = C.c_int(0); lh = C.c_int(0); no = C.c_int(0)
    torch.cuda.synchronize()
    _lib.check(lib.bm_sift_debug_level(_ptr(gray), h, w, octave, level, int(dog), _np_ptr(buf), C.byref(lw), C.byref(lh), C.byref(no)),
               "bm_sift_debug_level")
    return buf[:lw.value * lh.value].reshape(lh.value, lw.value).copy(), no.value
