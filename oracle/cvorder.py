"""cv2's keypoint ORDER, restated (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

The reference keeps whatever order `self.detector.detectAndCompute` returns (/root/reference/main.py:112,718) and that
order decides everything downstream: the match list (stable sort by distance, ties by query order, main.py:686-698), the
point order `cv2.findHomography` draws its cv::RNG subsets from (main.py:856-857) and therefore the homography itself.
Inside OpenCV 4.x the order is produced by `KeyPointsFilter::retainBest` (features2d/src/keypoint.cpp):

    std::nth_element(begin, begin + n - 1, end, ResponseGreater)     # libstdc++ introselect
    ambiguous = kp[n - 1].response
    new_end = std::partition(begin + n, end, response >= ambiguous)   # keep the ties
    resize(new_end)

ORB calls it per pyramid level on the FAST scores (row-major FAST output, 2 * quota) and again on the Harris responses
(quota); SIFT calls it once (700) after `removeDuplicatedSorted` has sorted the keypoints by (x, y, -size, angle, ...).
libstdc++ is a third-party dependency absent from /root/reference (it is compiled into the opencv-python wheel); the
restatement below follows bits/stl_algo.h / stl_heap.h (GCC 4.9 ... 14, unchanged across those releases) and is pinned
against (a) the real std:: algorithms (oracle/csrc/stl_order.cpp -> oracle/_ref/libstlorder.so) and (b) live cv2's ORB /
SIFT output order (tests/test_oracle_order_cpu.py).
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np


def _lg(n: int) -> int:
    return n.bit_length() - 1


def _adjust_heap(a, first, hole, length, value, gt):
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if gt(a[first + child], a[first + child - 1]):
            child -= 1
        a[first + hole] = a[first + child]
        hole = child
    if (length & 1) == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        a[first + hole] = a[first + child - 1]
        hole = child - 1
    parent = (hole - 1) // 2                                  # __push_heap
    while hole > top and gt(a[first + parent], value):
        a[first + hole] = a[first + parent]
        hole = parent
        parent = (hole - 1) // 2
    a[first + hole] = value


def _heap_select(a, first, middle, last, gt):
    length = middle - first
    if length >= 2:                                           # __make_heap
        parent = (length - 2) // 2
        while True:
            _adjust_heap(a, first, parent, length, a[first + parent], gt)
            if parent == 0:
                break
            parent -= 1
    for i in range(middle, last):
        if gt(a[i], a[first]):                                # __pop_heap(first, middle, i)
            value = a[i]
            a[i] = a[first]
            _adjust_heap(a, first, 0, length, value, gt)


def nth_element(a: list, nth: int, gt) -> None:
    """std::nth_element(a.begin(), a.begin() + nth, a.end(), gt) of libstdc++ (introselect), in place."""
    first, last = 0, len(a)
    if first == last or nth == last:
        return
    depth = _lg(last - first) * 2
    while last - first > 3:
        if depth == 0:
            _heap_select(a, first, nth + 1, last, gt)
            a[first], a[nth] = a[nth], a[first]
            return
        depth -= 1
        mid = first + (last - first) // 2
        x, y, z = first + 1, mid, last - 1                    # __move_median_to_first(first, first + 1, mid, last - 1)
        if gt(a[x], a[y]):
            pick = y if gt(a[y], a[z]) else (z if gt(a[x], a[z]) else x)
        else:
            pick = x if gt(a[x], a[z]) else (z if gt(a[y], a[z]) else y)
        a[first], a[pick] = a[pick], a[first]
        lo, hi, piv = first + 1, last, a[first]               # __unguarded_partition(first + 1, last, first)
        while True:
            while gt(a[lo], piv):
                lo += 1
            hi -= 1
            while gt(piv, a[hi]):
                hi -= 1
            if not lo < hi:
                break
            a[lo], a[hi] = a[hi], a[lo]
            lo += 1
        if lo <= nth:
            first = lo
        else:
            last = lo
    for i in range(first + 1, last):                          # __insertion_sort
        v = a[i]
        if gt(v, a[first]):
            a[first + 1:i + 1] = a[first:i]
            a[first] = v
        else:
            j = i
            while gt(v, a[j - 1]):
                a[j] = a[j - 1]
                j -= 1
            a[j] = v


def partition(a: list, lo: int, pred) -> int:
    """std::partition(a.begin() + lo, a.end(), pred) for bidirectional iterators; returns the new end."""
    first, last = lo, len(a)
    while True:
        while True:
            if first == last:
                return first
            if pred(a[first]):
                first += 1
            else:
                break
        last -= 1
        while True:
            if first == last:
                return first
            if not pred(a[last]):
                last -= 1
            else:
                break
        a[first], a[last] = a[last], a[first]
        first += 1


def retain_best_order(resp, n_points: int) -> np.ndarray:
    """Indices (into `resp`, input order) that KeyPointsFilter::retainBest keeps, in the order it leaves them."""
    r = [float(v) for v in np.asarray(resp, dtype=np.float32)]
    idx = list(range(len(r)))
    if n_points >= 0 and len(idx) > n_points:
        if n_points == 0:
            return np.zeros(0, np.int64)
        nth_element(idx, n_points - 1, lambda i, j: r[i] > r[j])
        amb = r[idx[n_points - 1]]
        idx = idx[:partition(idx, n_points, lambda i: r[i] >= amb)]
    return np.asarray(idx, dtype=np.int64)


def nth_element_order(resp, nth: int) -> np.ndarray:
    r = [float(v) for v in np.asarray(resp, dtype=np.float32)]
    idx = list(range(len(r)))
    nth_element(idx, nth, lambda i, j: r[i] > r[j])
    return np.asarray(idx, dtype=np.int64)


def adversarial_input(n: int, nth: int) -> np.ndarray:
    """McIlroy's adversary ("A Killer Adversary for Quicksort", 1999) run against the restated introselect: float32 values that
    drive std::nth_element(.., nth, ..) into its heap-select fallback.  Used by the tests of the depth-limit path."""
    GAS = 1 << 30
    val = [GAS] * n
    state = {"nsolid": 0, "cand": 0}

    def gt(x, y):
        if val[x] == GAS and val[y] == GAS:
            f = x if x == state["cand"] else y
            val[f] = state["nsolid"]
            state["nsolid"] += 1
        if val[x] == GAS:
            state["cand"] = x
        elif val[y] == GAS:
            state["cand"] = y
        return val[x] > val[y]
    idx = list(range(n))
    nth_element(idx, nth, gt)
    for i in range(n):
        if val[i] == GAS:
            val[i] = state["nsolid"]
            state["nsolid"] += 1
    return np.asarray(val, np.float32)


# ---- the real libstdc++ (oracle/_ref/libstlorder.so), when built ----------------------------------------------------
_REF = Path(__file__).resolve().parent / "_ref" / "libstlorder.so"
_lib = None


def stl_available() -> bool:
    return _REF.exists()


def _stl():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(_REF))
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        _lib.stl_retain_best.argtypes = [fp, ctypes.c_int, ctypes.c_int, ip]
        _lib.stl_retain_best.restype = ctypes.c_int
        _lib.stl_nth_element.argtypes = [fp, ctypes.c_int, ctypes.c_int, ip]
        _lib.stl_sift_sort_unique.argtypes = [fp, ctypes.c_int, ip]
        _lib.stl_sift_sort_unique.restype = ctypes.c_int
    return _lib


def stl_retain_best(resp, n_points: int) -> np.ndarray:
    r = np.ascontiguousarray(resp, dtype=np.float32)
    out = np.zeros(max(len(r), 1), np.int32)
    m = _stl().stl_retain_best(r.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(r), int(n_points),
                               out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return out[:m].astype(np.int64)


def stl_nth_element(resp, nth: int) -> np.ndarray:
    r = np.ascontiguousarray(resp, dtype=np.float32)
    out = np.zeros(max(len(r), 1), np.int32)
    _stl().stl_nth_element(r.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(r), int(nth),
                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return out[:len(r)].astype(np.int64)


def stl_sift_sort_unique(kp7) -> np.ndarray:
    k = np.ascontiguousarray(kp7, dtype=np.float32)
    out = np.zeros(max(len(k), 1), np.int32)
    m = _stl().stl_sift_sort_unique(k.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(k),
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return out[:m].astype(np.int64)
