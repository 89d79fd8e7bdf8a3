"""ctypes binding of libb200mosaic.so (the C ABI in include/b200mosaic.h).  No CPU fallback: if the library is
missing or CUDA is unavailable every call raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libb200mosaic.so"

BM_OK, BM_SKIP_FEW_MATCHES, BM_SKIP_NO_H, BM_REJECTED_IDENTITY = 0, 1, 2, 3
BM_DET_SIFT, BM_DET_ORB = 0, 1
BM_VAL_OK, BM_VAL_NAN, BM_VAL_TRANSLATION, BM_VAL_SCALE, BM_VAL_PERSPECTIVE = range(5)


class BmConfig(C.Structure):
    _fields_ = [("frame_h", C.c_int), ("frame_w", C.c_int), ("canvas_h", C.c_int), ("canvas_w", C.c_int),
                ("detector", C.c_int), ("nfeatures", C.c_int), ("device", C.c_int)]


class BmFrameInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("n_kp_cur", C.c_int), ("n_kp_prev", C.c_int), ("n_matches", C.c_int),
                ("ransac_iters", C.c_int), ("n_inliers", C.c_int), ("validate_reason", C.c_int),
                ("any_overlap", C.c_int), ("win", C.c_int * 4), ("validate_value", C.c_double),
                ("H_rel", C.c_double * 9), ("H", C.c_double * 9)]


class B200MosaicError(RuntimeError):
    pass


_lib = None


def _sig(lib, name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


def load():
    """Load the shared library (building it first if it is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    # build.build() is digest-stamped: it returns at once when lib/ matches the sources and rebuilds a stale binary otherwise
    # (the directory is git-ignored but travels with the tree).  Without the sources next to it the library is used as is.
    from . import build as _build
    if not LIB_PATH.exists() or (_build.CSRC.exists() and any(_build.CSRC.glob("*.cu")) and os.path.exists(_build.NVCC)):
        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    vp, i, sz, dp = C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    _sig(lib, "bm_last_error", C.c_char_p)
    _sig(lib, "bm_version", i)
    _sig(lib, "bm_create", i, C.POINTER(BmConfig), C.POINTER(vp))
    _sig(lib, "bm_destroy", i, vp)
    _sig(lib, "bm_first_frame", i, vp, vp, sz)
    _sig(lib, "bm_process_frame", i, vp, vp, sz, C.POINTER(BmFrameInfo))
    _sig(lib, "bm_get_canvas", i, vp, vp)
    _sig(lib, "bm_set_canvas", i, vp, vp)
    _sig(lib, "bm_get_state", i, vp, dp, ip, dp)
    _sig(lib, "bm_set_stabilization", i, vp, i, i, C.c_double, C.c_double)
    _sig(lib, "bm_validate_homography", i, dp, C.c_double, C.c_double, dp)
    _sig(lib, "bm_alloc_pinned", i, sz, C.POINTER(vp))
    _sig(lib, "bm_free_pinned", i, vp)
    _sig(lib, "bm_warp_frame", i, vp, vp, sz, dp, C.POINTER(BmFrameInfo))
    _sig(lib, "bm_warp_frame_device", i, vp, vp, dp, C.POINTER(BmFrameInfo))
    _sig(lib, "bm_sync", i, vp)
    _sig(lib, "bm_process_frame_device", i, vp, vp, C.POINTER(BmFrameInfo))
    _sig(lib, "bm_process_frame_begin", i, vp, vp, sz)
    _sig(lib, "bm_process_frame_begin_device", i, vp, vp)
    _sig(lib, "bm_process_frame_end", i, vp, C.POINTER(BmFrameInfo))
    _sig(lib, "bm_estimate_frame", i, vp, vp, sz, vp, C.POINTER(BmFrameInfo))
    _sig(lib, "bm_prefetch_frame", i, vp, vp, sz)
    _sig(lib, "bm_prefetch_frame_device", i, vp, vp)
    _sig(lib, "bm_finalize", i, vp, i, i, i, i, vp, sz, ip, ip)
    _sig(lib, "bm_preview", i, vp, i, i, i, vp, sz)
    _sig(lib, "bm_jpeg_bound", sz, i, i)
    _sig(lib, "bm_jpeg_encode", i, vp, i, i, i, i, vp, sz, C.POINTER(sz))
    _sig(lib, "bm_finalize_jpeg", i, vp, i, i, i, i, i, vp, sz, C.POINTER(sz), ip, ip)
    _sig(lib, "bm_warm_up", i, vp)
    _sig(lib, "bm_warp_frame_async", i, vp, vp, sz, dp)
    _sig(lib, "bm_set_overlap", i, vp, i)
    _sig(lib, "bm_clear_canvas", i, vp)
    _sig(lib, "bm_get_canvas_device", i, vp, vp)
    _sig(lib, "bm_tile_set_ghost", i, vp, i, vp)
    _sig(lib, "bm_tile_export_carries", i, vp, i, i, vp)
    _sig(lib, "bm_tile_export_rect", i, vp, i, i, i, i, vp)
    _sig(lib, "bm_tile_import_rect", i, vp, i, i, i, i, vp)
    _sig(lib, "bm_timing_enable", i, vp, i)
    _sig(lib, "bm_timing_read", i, vp, dp, dp, ip, i)
    _sig(lib, "bm_kernel_launches", C.c_longlong)
    _sig(lib, "bm_stream", vp, vp)
    _sig(lib, "bm_upload_frame", i, vp, vp, sz, C.POINTER(vp))
    _sig(lib, "bm_ingest_bgr", i, vp, i, i, vp, vp, vp)
    _sig(lib, "bm_warp_perspective_bgr", i, vp, i, i, dp, vp, i, i, vp)
    _sig(lib, "bm_distance_transform", i, vp, i, i, vp, vp)
    _sig(lib, "bm_gaussian_blur31", i, vp, i, i, vp, vp)
    _sig(lib, "bm_blend_step_bgr", i, vp, vp, i, i, ip, ip, vp)
    fp = C.POINTER(C.c_float)
    _sig(lib, "bm_orb_detect_and_compute", i, vp, i, i, i, vp, vp, i, ip)
    _sig(lib, "bm_sift_detect_and_compute", i, vp, i, i, i, vp, vp, i, ip)
    _sig(lib, "bm_cv_retain_best", i, vp, i, i, i, vp, ip)
    _sig(lib, "bm_match_hamming_crosscheck", i, vp, i, vp, i, vp, vp, vp, ip)
    _sig(lib, "bm_match_l2_knn2_ratio", i, vp, i, vp, i, C.c_double, vp, vp, vp, ip)
    _sig(lib, "bm_ransac_homography", i, vp, vp, i, C.c_double, i, C.c_double, dp, ip, ip, ip)
    _sig(lib, "bm_ransac_profile", i, vp, vp, i, C.c_double, i, C.c_double, dp, vp, ip, ip)
    _sig(lib, "bm_debug_lm_force_eig", i, i)
    _sig(lib, "bm_debug_lm_stats", i, vp, i)
    _sig(lib, "bm_get_keypoints", i, vp, i, vp, vp, i, ip)
    _sig(lib, "bm_get_matches", i, vp, vp, vp, vp, i, ip)
    _sig(lib, "bm_keypoint_capacity", i)
    _sig(lib, "bm_orb_debug_level", i, vp, i, i, i, vp, vp, ip, ip)
    _sig(lib, "bm_sift_pyramid_ms", i, vp, i, i, i, dp, dp)
    _sig(lib, "bm_match_l2_ms", i, i, i, i, dp, dp)
    _sig(lib, "bm_sift_debug_level", i, vp, i, i, i, i, i, vp, ip, ip, ip)
    _lib = lib
    return lib


def check(status: int, what: str = "") -> int:
    if status < 0:
        msg = load().bm_last_error().decode("utf-8", "replace")
        raise B200MosaicError(f"{what or 'libb200mosaic'} failed ({status}): {msg}")
    return status


def validate_homography(H, translation_threshold=50.0, scale_threshold=0.3):
    """main.py:761-801 through the one native implementation.  Returns (BM_VAL_* reason, value)."""
    if H is None:
        return BM_VAL_NAN, 0.0
    _a, hp = dbl9(H)
    v = C.c_double(0.0)
    r = load().bm_validate_homography(hp, float(translation_threshold), float(scale_threshold), C.byref(v))
    return r, v.value


def dbl9(H):
    import numpy as np
    a = np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(9))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))
