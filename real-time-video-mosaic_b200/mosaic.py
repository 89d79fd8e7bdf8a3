"""VideMosaic -- host-side mirror of the reference class (`/root/reference/main.py:15-112, 676-977`) over the C ABI.

Same constructor, methods, attributes, prints and soft-failure behaviour as the reference for the stitching path;
all pixel / feature arithmetic runs in libb200mosaic.so (hand-written sm_100a CUDA).  Out-of-scope methods of the
reference class (`detect_objects`, ... -- YOLO, main.py:114-674) are delegated untouched to the reference class when
one is registered with `VideMosaic.reference_class = main.VideMosaic` (see INTEGRATION.md / run.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class _KeyPoint:
    """cv2.KeyPoint look-alike (attributes the reference reads: .pt; plus the rest of the cv2 fields)."""
    __slots__ = ("pt", "size", "angle", "response", "octave", "class_id")

    def __init__(self, x, y, size=0.0, angle=-1.0, response=0.0, octave=0, class_id=-1):
        self.pt = (float(x), float(y))
        self.size, self.angle, self.response, self.octave, self.class_id = size, angle, response, octave, class_id


class DMatch:
    """cv2.DMatch look-alike (queryIdx, trainIdx, distance; imgIdx = 0)."""
    __slots__ = ("queryIdx", "trainIdx", "distance", "imgIdx")

    def __init__(self, q, t, d):
        self.queryIdx, self.trainIdx, self.distance, self.imgIdx = int(q), int(t), float(d), 0

    def __repr__(self):
        return f"DMatch(q={self.queryIdx}, t={self.trainIdx}, d={self.distance:g})"


class VideMosaic:
    reference_class = None      # set to the reference's own class to delegate detect_objects & co.

    def __init__(self, first_image, output_height_times=2, output_width_times=1.2, detector_type="sift",
                 show_intermediate=True, output_dir=None, visualize=True, *, canvas_size=None, device=0,
                 nfeatures=700):
        """Same positional/keyword arguments as main.py:17.  Extra keyword-only arguments: `canvas_size=(Hc,Wc)`
        (explicit canvas, e.g. 32768x32768 of config 5, not expressible through the float multipliers), `device`."""
        self.detector_type = detector_type
        self.show_intermediate = show_intermediate
        self.output_dir = output_dir
        self.visualize = visualize
        if detector_type not in ("sift", "orb"):
            # the reference leaves self.detector undefined and fails later with AttributeError (main.py:32-37)
            raise AttributeError("'VideMosaic' object has no attribute 'detector'")
        first_image = self._check_frame(first_image)
        fh, fw, fc = first_image.shape
        if canvas_size is None:
            ch, cw = int(output_height_times * fh), int(output_width_times * fw)      # main.py:80-81
        else:
            ch, cw = int(canvas_size[0]), int(canvas_size[1])
        self._lib = _lib.load()
        cfg = _lib.BmConfig(frame_h=fh, frame_w=fw, canvas_h=ch, canvas_w=cw,
                            detector=_lib.BM_DET_SIFT if detector_type == "sift" else _lib.BM_DET_ORB,
                            nfeatures=nfeatures, device=device)
        self._h = C.c_void_p()
        _lib.check(self._lib.bm_create(C.byref(cfg), C.byref(self._h)), "bm_create")
        self._shape = (ch, cw, fc)
        self._frame_shape = (fh, fw, fc)
        self._canvas_cache = None
        self.frame_prev = first_image
        _lib.check(self._lib.bm_first_frame(self._h, first_image.ctypes.data_as(C.c_void_p), 0), "bm_first_frame")
        self.w_offset = int(ch / 1 - fh / 1)                      # main.py:86 (row offset, names swapped upstream)
        self.h_offset = int(cw / 2 - fw / 2)                      # main.py:87
        self.H_old = np.eye(3)
        self.H_old[0, 2] = self.h_offset
        self.H_old[1, 2] = self.w_offset
        self.H = None
        self.stabilization_enabled = True                          # main.py:97-102
        self.homography_history = []
        self.history_size = 5
        self.translation_threshold = 50
        self.scale_threshold = 0.3
        self.last_valid_H = np.eye(3)
        self.last_info = None
        self._ref_delegate = None

    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _check_frame(frame):
        if not isinstance(frame, np.ndarray) or frame.ndim != 3 or frame.shape[2] != 3 or frame.dtype != np.uint8:
            raise TypeError("frame must be an (H, W, 3) uint8 BGR ndarray (what cv2.VideoCapture.read returns)")
        return np.ascontiguousarray(frame)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._lib.bm_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    close = __del__

    # ---- output_img: lazy device->host copy, cached until the next frame -----------------------
    @property
    def output_img(self):
        """The canvas as a uint8 (Hc, Wc, 3) array (the reference holds float64 that only ever contains integers; every caller
        casts or copies).  The array is a host COPY of the device canvas, cached until the next frame: writes into it
        do not reach the device -- assign a whole image (`vm.output_img = img`) to replace the canvas."""
        if self._canvas_cache is None:
            out = np.empty(self._shape, dtype=np.uint8)
            _lib.check(self._lib.bm_get_canvas(self._h, out.ctypes.data_as(C.c_void_p)), "bm_get_canvas")
            self._canvas_cache = out
        return self._canvas_cache

    @output_img.setter
    def output_img(self, img):
        """`vm.output_img = array` as reference callers may do: replaces the device canvas (values are truncated to uint8 like
        every consumer of the reference's float64 canvas does)."""
        a = np.ascontiguousarray(np.asarray(img).astype(np.uint8))
        if tuple(a.shape) != tuple(self._shape):
            raise ValueError(f"output_img must have shape {tuple(self._shape)}")
        _lib.check(self._lib.bm_set_canvas(self._h, a.ctypes.data_as(C.c_void_p)), "bm_set_canvas")
        self._canvas_cache = None

    def read_canvas(self, out):
        """canvas into a caller-owned uint8 (Hc, Wc, 3) C-contiguous array (e.g. a pinned buffer that is reused for every fetch:
        the copy then runs at PCIe speed and no fresh 15 MB array is page-faulted in).  Returns `out`."""
        if out.dtype != np.uint8 or tuple(out.shape) != tuple(self._shape) or not out.flags["C_CONTIGUOUS"]:
            raise ValueError(f"read_canvas needs a C-contiguous uint8 array of shape {tuple(self._shape)}")
        _lib.check(self._lib.bm_get_canvas(self._h, out.ctypes.data_as(C.c_void_p)), "bm_get_canvas")
        return out

    # ---- main.py:710-759 -------------------------------------------------------------------------------------
    def process_frame(self, frame_cur, frame_count=0, next_frame=None, next2_frame=None, next3_frame=None):
        """main.py:710-759.  `next_frame` / `next2_frame` / `next3_frame` (optional, not in the reference): the frames of the NEXT
        calls, if the caller already has them (up to LOOKAHEAD = 3) -- their H2D copies and detectAndCompute then overlap this frame's
        processing (multi-buffered ingest, detect-ahead); pass the same array objects to the following process_frame calls."""
        frame_cur = self._check_frame(frame_cur)
        if frame_cur.shape != self._frame_shape:
            raise ValueError(f"frame shape {frame_cur.shape} != first frame shape {self._frame_shape}")
        self.frame_cur = frame_cur
        self._sync_knobs()
        info = _lib.BmFrameInfo()
        if next_frame is None:
            st = _lib.check(self._lib.bm_process_frame(self._h, frame_cur.ctypes.data_as(C.c_void_p), 0, C.byref(info)),
                            "bm_process_frame")
        else:
            next_frame = self._check_frame(next_frame)
            _lib.check(self._lib.bm_process_frame_begin(self._h, frame_cur.ctypes.data_as(C.c_void_p), 0), "bm_process_frame_begin")
            _lib.check(self._lib.bm_prefetch_frame(self._h, next_frame.ctypes.data_as(C.c_void_p), 0), "bm_prefetch_frame")
            self._next_frame_ref = [next_frame]                    # keep the staged buffers alive / unchanged until they are consumed
            for more in (next2_frame, next3_frame):
                if more is None:
                    break
                more = self._check_frame(more)
                _lib.check(self._lib.bm_prefetch_frame(self._h, more.ctypes.data_as(C.c_void_p), 0), "bm_prefetch_frame")
                self._next_frame_ref.append(more)
            st = _lib.check(self._lib.bm_process_frame_end(self._h, C.byref(info)), "bm_process_frame_end")
        self.last_info = info
        if st == _lib.BM_SKIP_FEW_MATCHES:
            print(f"Предупреждение: Недостаточно совпадений ({info.n_matches}), пропуск кадра")          # :723
            return
        if st == _lib.BM_SKIP_NO_H:
            print("Предупреждение: Не удалось вычислить гомографию, пропуск кадра")                      # :730
            return
        H_rel = np.array(info.H_rel, dtype=np.float64).reshape(3, 3)
        if st == _lib.BM_REJECTED_IDENTITY:
            self._print_validate(info.validate_reason, info.validate_value)
            print("Предупреждение: Невалидная гомография (тряска/размытие), использую последнюю валидную")  # :735
            H_used = np.eye(3)
        else:
            self.last_valid_H = H_rel.copy()
            H_used = H_rel
        if self.stabilization_enabled:                             # smooth_homography returns before appending otherwise (:812-816)
            self.homography_history.append(H_used.copy())
            if len(self.homography_history) > self.history_size:
                self.homography_history.pop(0)
        self.H = np.array(info.H, dtype=np.float64).reshape(3, 3)
        self.H_old = self.H
        self.frame_prev = frame_cur
        self._canvas_cache = None

    # ---- device-resident / raw-pointer variants used by bench.py ------------------------------------------------
    def process_frame_ptr(self, host_ptr, next_ptr=None, next2_ptr=None, next3_ptr=None):
        """process_frame on a raw host pointer (e.g. pinned memory): no NumPy checks, no prints.  Returns the status.
        next_ptr / next2_ptr / next3_ptr: host pointers of the next frames, staged (H2D + ingest on the copy stream, detect-ahead) while
        this one is processed."""
        info = _lib.BmFrameInfo()
        if next_ptr is None:
            st = _lib.check(self._lib.bm_process_frame(self._h, C.c_void_p(host_ptr), 0, C.byref(info)), "bm_process_frame")
        else:
            _lib.check(self._lib.bm_process_frame_begin(self._h, C.c_void_p(host_ptr), 0), "bm_process_frame_begin")
            _lib.check(self._lib.bm_prefetch_frame(self._h, C.c_void_p(next_ptr), 0), "bm_prefetch_frame")
            for more in (next2_ptr, next3_ptr):
                if more is None:
                    break
                _lib.check(self._lib.bm_prefetch_frame(self._h, C.c_void_p(more), 0), "bm_prefetch_frame")
            st = _lib.check(self._lib.bm_process_frame_end(self._h, C.byref(info)), "bm_process_frame_end")
        self.last_info = info
        self._canvas_cache = None
        return st

    def process_frame_device(self, dev_ptr, next_ptr=None, next2_ptr=None, next3_ptr=None):
        """process_frame on a frame already in device memory (packed BGR).  Returns the status.  next_ptr: device pointer of the
        next frame (ingested and started while this one is finished, like process_frame_ptr's next_ptr)."""
        info = _lib.BmFrameInfo()
        if next_ptr is None:
            st = _lib.check(self._lib.bm_process_frame_device(self._h, C.c_void_p(dev_ptr), C.byref(info)), "bm_process_frame_device")
        else:
            _lib.check(self._lib.bm_process_frame_begin_device(self._h, C.c_void_p(dev_ptr)), "bm_process_frame_begin_device")
            _lib.check(self._lib.bm_prefetch_frame_device(self._h, C.c_void_p(next_ptr)), "bm_prefetch_frame_device")
            for more in (next2_ptr, next3_ptr):
                if more is None:
                    break
                _lib.check(self._lib.bm_prefetch_frame_device(self._h, C.c_void_p(more)), "bm_prefetch_frame_device")
            st = _lib.check(self._lib.bm_process_frame_end(self._h, C.byref(info)), "bm_process_frame_end")
        self.last_info = info
        self._canvas_cache = None
        return st

    def begin_frame_ptr(self, host_ptr):
        """enqueue half of process_frame (H2D + detect + match + RANSAC), no wait -- see bm_process_frame_begin"""
        _lib.check(self._lib.bm_process_frame_begin(self._h, C.c_void_p(host_ptr), 0), "bm_process_frame_begin")

    def prefetch_ptr(self, host_ptr):
        """stage a frame the caller will process soon (H2D + ingest now, detect-ahead during the next end_frame) -- bm_prefetch_frame;
        up to three frames ahead, in processing order"""
        _lib.check(self._lib.bm_prefetch_frame(self._h, C.c_void_p(host_ptr), 0), "bm_prefetch_frame")

    def begin_frame_device(self, dev_ptr):
        _lib.check(self._lib.bm_process_frame_begin_device(self._h, C.c_void_p(dev_ptr)), "bm_process_frame_begin_device")

    def end_frame(self):
        """second half: wait for the read-back, host control flow, enqueue warp/blend.  Returns the status."""
        info = _lib.BmFrameInfo()
        st = _lib.check(self._lib.bm_process_frame_end(self._h, C.byref(info)), "bm_process_frame_end")
        self.last_info = info
        self._canvas_cache = None
        return st

    def estimate_frame(self, frame, next_frame=None, later_frames=()):
        """offline pair mode: features + matches + RANSAC against the previous frame, no validation / warp; the frame
        becomes the new previous.  Returns (status, H_rel or None, n_matches).  next_frame: the frame of the next call (its
        upload overlaps this pair's estimation; pass the same array object next time).  later_frames: up to two frames after that one,
        staged as well (their uploads start now, their features are computed during the next call)."""
        frame = self._check_frame(frame)
        nxt = None
        keep = []
        if next_frame is not None:
            keep.append(self._check_frame(next_frame))
            nxt = keep[0].ctypes.data_as(C.c_void_p)
        info = _lib.BmFrameInfo()
        st = _lib.check(self._lib.bm_estimate_frame(self._h, frame.ctypes.data_as(C.c_void_p), 0, nxt, C.byref(info)), "bm_estimate_frame")
        if nxt is not None:
            for f in later_frames:
                f = self._check_frame(f)
                _lib.check(self._lib.bm_prefetch_frame(self._h, f.ctypes.data_as(C.c_void_p), 0), "bm_prefetch_frame")
                keep.append(f)
        self._next_frame_ref = keep                                # the staged buffers stay alive / unchanged until they are consumed
        self.last_info = info
        H = np.array(info.H_rel, dtype=np.float64).reshape(3, 3) if st == _lib.BM_OK else None
        return st, H, info.n_matches

    def finalize(self, threshold=80, margin=30, target_w=None, target_h=None):
        """crop_black_areas(output_img, threshold, margin) + scale_to_screen(cropped, target_w, target_h) (main.py:980-1038) on the
        device canvas, as main() does before cv2.imwrite('mosaic.jpg') (main.py:1647-1659): only the screen-sized result is copied
        to the host.  Returns the uint8 BGR image; `self.last_crop_rect` = (x, y, w, h) of the crop."""
        wh = (C.c_int * 2)(); rect = (C.c_int * 4)()
        tw, th = (int(target_w), int(target_h)) if target_w and target_h else (0, 0)
        _lib.check(self._lib.bm_finalize(self._h, int(threshold), int(margin), tw, th, None, 0, wh, rect), "bm_finalize")
        out = np.empty((wh[1], wh[0], 3), dtype=np.uint8)
        _lib.check(self._lib.bm_finalize(self._h, int(threshold), int(margin), tw, th, out.ctypes.data_as(C.c_void_p), out.nbytes, wh, rect),
                   "bm_finalize")
        self.last_crop_rect = tuple(rect)
        return out

    def finalize_jpeg(self, threshold=80, margin=30, target_w=None, target_h=None, quality=95):
        """main.py:1647-1666 in one device pass: crop_black_areas + scale_to_screen + the bytes cv2.imwrite('mosaic.jpg', scaled) writes
        (baseline JPEG, default quality 95; byte-identical to cv2's file).  Only the compressed file is copied to the host.
        Returns `bytes`; `self.last_crop_rect`, `self.last_final_size` = (w, h) describe the image inside."""
        wh = (C.c_int * 2)(); rect = (C.c_int * 4)(); n = C.c_size_t(0)
        tw, th = (int(target_w), int(target_h)) if target_w and target_h else (0, 0)
        _lib.check(self._lib.bm_finalize(self._h, int(threshold), int(margin), tw, th, None, 0, wh, rect), "bm_finalize")
        out = np.empty(self._lib.bm_jpeg_bound(wh[0], wh[1]), dtype=np.uint8)
        _lib.check(self._lib.bm_finalize_jpeg(self._h, int(threshold), int(margin), tw, th, int(quality), out.ctypes.data_as(C.c_void_p),
                                              out.nbytes, C.byref(n), wh, rect), "bm_finalize_jpeg")
        self.last_crop_rect = tuple(rect)
        self.last_final_size = (wh[0], wh[1])
        return out[:n.value].tobytes()

    def preview(self, size=(400, 300), rgb=True):
        """Thumbnail of the live canvas made on the device: bit-identical to what the GUI computes from `output_img.copy()`
        (main.py:1630-1632 -> gui.py:143-158: cv2.cvtColor(BGR2RGB), Image.fromarray(...).resize(size), Pillow's default bicubic),
        but only size[0] * size[1] * 3 bytes are copied to the host.  Returns uint8 (size[1], size[0], 3), RGB (or BGR)."""
        w, h = int(size[0]), int(size[1])
        out = np.empty((h, w, 3), dtype=np.uint8)
        _lib.check(self._lib.bm_preview(self._h, w, h, 1 if rgb else 0, out.ctypes.data_as(C.c_void_p), out.nbytes), "bm_preview")
        return out

    def warm_up(self):
        """capture every CUDA graph of the per-frame path now (optional; avoids capture hiccups in the first frames)"""
        _lib.check(self._lib.bm_warm_up(self._h), "bm_warm_up")

    def set_overlap(self, on):
        """True (default): the warp/blend chain of frame t overlaps detect/match/RANSAC of frame t+1; False: strictly serial"""
        _lib.check(self._lib.bm_set_overlap(self._h, 1 if on else 0), "bm_set_overlap")

    def clear_canvas(self):
        _lib.check(self._lib.bm_clear_canvas(self._h), "bm_clear_canvas")
        self._canvas_cache = None

    def canvas_to_device(self, dev_ptr):
        """packed BGR canvas into a device buffer (torch tensor .data_ptr()) -- for NCCL gathers of canvas tiles"""
        _lib.check(self._lib.bm_get_canvas_device(self._h, C.c_void_p(dev_ptr)), "bm_get_canvas_device")

    def timing(self, enable=None, reset=False):
        """CUDA-event timing of the warp/blend chain: returns (ms, algorithmic bytes = 3N + 6A per frame, frames)."""
        if enable is not None:
            _lib.check(self._lib.bm_timing_enable(self._h, int(enable)))
        ms = C.c_double(0); by = C.c_double(0); fr = C.c_int(0)
        _lib.check(self._lib.bm_timing_read(self._h, C.byref(ms), C.byref(by), C.byref(fr), int(reset)))
        return ms.value, by.value, fr.value

    def sync(self):
        _lib.check(self._lib.bm_sync(self._h))

    # ---- features / matches of the last frame (lazy small D2H), cv2-like objects --------------------------------
    def _fetch_kp(self, which):
        cap = self._lib.bm_keypoint_capacity()
        dbytes = 32 if self.detector_type == "orb" else 128
        kp = np.empty((cap, 6), np.float32); des = np.empty((cap, dbytes), np.uint8); n = C.c_int(0)
        _lib.check(self._lib.bm_get_keypoints(self._h, which, kp.ctypes.data_as(C.c_void_p), des.ctypes.data_as(C.c_void_p),
                                              cap, C.byref(n)), "bm_get_keypoints")
        kps = [_KeyPoint(r[0], r[1], float(r[2]), float(r[3]), float(r[4]), int(r[5])) for r in kp[:n.value]]
        d = des[:n.value].copy()
        return kps, (d if self.detector_type == "orb" else d.astype(np.float32))

    @property
    def kp_prev(self):
        return self._fetch_kp(0)[0]

    @property
    def des_prev(self):
        return self._fetch_kp(0)[1]

    @property
    def kp_cur(self):
        """keypoints of the last processed frame (main.py:718).  After an accepted frame they are also `kp_prev` (main.py:757)."""
        return self._fetch_kp(1)[0]

    @property
    def des_cur(self):
        return self._fetch_kp(1)[1]

    @property
    def matches(self):
        cap = self._lib.bm_keypoint_capacity()
        q = np.empty(cap, np.int32); t = np.empty(cap, np.int32); d = np.empty(cap, np.float32); m = C.c_int(0)
        _lib.check(self._lib.bm_get_matches(self._h, q.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),
                                            d.ctypes.data_as(C.c_void_p), cap, C.byref(m)), "bm_get_matches")
        return [DMatch(q[i], t[i], d[i]) for i in range(m.value)]

    # ---- main.py:676-708 ---------------------------------------------------------------------------------------
    def match(self, des_cur, des_prev):
        from . import ops
        if self.detector_type == "sift":
            m = ops.match_l2_ratio(np.asarray(des_cur), np.asarray(des_prev), 0.7)
        else:
            m = ops.match_hamming_crosscheck(np.asarray(des_cur), np.asarray(des_prev))
        return [DMatch(r[0], r[1], r[2]) for r in m]

    # ---- main.py:836-859 ---------------------------------------------------------------------------------------
    @staticmethod
    def findHomography(image_1_kp, image_2_kp, matches):
        from . import ops
        p1 = np.zeros((len(matches), 2), dtype=np.float32)
        p2 = np.zeros((len(matches), 2), dtype=np.float32)
        for i, m in enumerate(matches):
            p1[i] = image_1_kp[m.queryIdx].pt
            p2[i] = image_2_kp[m.trainIdx].pt
        H, _, _ = ops.ransac_homography(p1, p2, 2.0)
        return H

    def _sync_knobs(self):
        _lib.check(self._lib.bm_set_stabilization(self._h, int(bool(self.stabilization_enabled)), int(self.history_size),
                                                  float(self.translation_threshold), float(self.scale_threshold)))

    @staticmethod
    def _print_validate(reason, value):
        if reason == _lib.BM_VAL_TRANSLATION:
            print(f"Предупреждение: Обнаружено большое смещение ({value:.1f}px), возможна тряска")        # :788
        elif reason == _lib.BM_VAL_SCALE:
            print(f"Предупреждение: Обнаружено большое изменение масштаба ({value:.2f}), возможна тряска")  # :793
        elif reason == _lib.BM_VAL_PERSPECTIVE:
            print("Предупреждение: Обнаружены сильные перспективные искажения")                           # :798

    # ---- main.py:761-801 (NumPy mirror for direct callers; the frame loop uses the native copy) -------------
    def validate_homography(self, H):
        if H is None:
            return False
        reason, value = _lib.validate_homography(H, self.translation_threshold, self.scale_threshold)
        self._print_validate(reason, value)
        return reason == _lib.BM_VAL_OK

    # ---- main.py:803-834 ---------------------------------------------------------------------------------------
    def smooth_homography(self, H):
        if not self.stabilization_enabled:
            return H
        self.homography_history.append(H.copy())
        if len(self.homography_history) > self.history_size:
            self.homography_history.pop(0)
        if len(self.homography_history) < 2:
            return H
        weights = np.linspace(0.5, 1.0, len(self.homography_history))
        weights = weights / np.sum(weights)
        out = np.zeros_like(H)
        for w, h in zip(weights, self.homography_history):
            out += w * h
        return out

    # ---- main.py:861-936 ---------------------------------------------------------------------------------------
    def warp(self, frame_cur, H):
        frame_cur = self._check_frame(frame_cur)
        _a, hp = _lib.dbl9(H)
        info = _lib.BmFrameInfo()
        _lib.check(self._lib.bm_warp_frame(self._h, frame_cur.ctypes.data_as(C.c_void_p), 0, hp, C.byref(info)),
                   "bm_warp_frame")
        self.last_info = info
        self._canvas_cache = None
        if self.visualize:                                         # display tail, main.py:929-934 (host, optional)
            import cv2
            tmp = self.draw_border(np.copy(self.output_img), self.get_transformed_corners(frame_cur, np.asarray(H)),
                                   color=(0, 0, 255))
            cv2.namedWindow('output', cv2.WINDOW_NORMAL)
            cv2.imshow('output', tmp / 255.)
        return self.output_img

    def warp_nosync(self, frame_cur, H):
        """warp(frame, H) without the display tail and without waiting for the device (canvas-tile mode)"""
        frame_cur = self._check_frame(frame_cur)
        _a, hp = _lib.dbl9(H)
        self._keep = frame_cur                                     # pageable frames are staged synchronously; keep pinned ones alive
        _lib.check(self._lib.bm_warp_frame_async(self._h, frame_cur.ctypes.data_as(C.c_void_p), 0, hp), "bm_warp_frame_async")
        self._canvas_cache = None

    @staticmethod
    def get_transformed_corners(frame_cur, H):                     # main.py:938-962 (4 points, host)
        h, w = frame_cur.shape[:2]
        pts = np.array([[0, 0, 1], [w, 0, 1], [w, h, 1], [0, h, 1]], dtype=np.float64).T
        q = np.asarray(H, dtype=np.float64) @ pts
        q = (q[:2] / q[2]).T.astype(np.float32)
        return np.array(q[None], dtype=np.int32)

    def draw_border(self, image, corners, color=(0, 0, 0)):       # main.py:964-977 (display only)
        import cv2
        for i in range(corners.shape[1] - 1, -1, -1):
            cv2.line(image, tuple(int(v) for v in corners[0, i, :]), tuple(int(v) for v in corners[0, i - 1, :]),
                     thickness=5, color=color)
        return image

    # ---- out-of-scope methods stay on the reference's path -----------------------------------------------------
    def __getattr__(self, name):
        # only called for attributes that are not found normally
        if name.startswith("_") or VideMosaic.reference_class is None:
            raise AttributeError(f"'VideMosaic' object has no attribute '{name}'")
        d = self.__dict__.get("_ref_delegate")
        if d is None:
            d = VideMosaic.reference_class(np.ascontiguousarray(self.frame_prev[:64, :64]), detector_type="orb",
                                           show_intermediate=False, output_dir=self.output_dir, visualize=False)
            self.__dict__["_ref_delegate"] = d
        return getattr(d, name)
