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
