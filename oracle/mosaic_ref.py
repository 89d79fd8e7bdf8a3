"""RefMosaic -- CPU restatement of the reference stitcher (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows `/root/reference/main.py` `class VideMosaic` for the hot path only:
ctor :17-102, process_first_frame :104-112, match :676-708, process_frame :710-759,
validate_homography :761-801, smooth_homography :803-834, findHomography :836-859, warp :861-936.
It performs the same cv2 / NumPy calls with the same dtypes and the same control flow, so for identical
inputs it is bit-identical to the unmodified reference (pinned by tests/golden fixtures).  The YOLO /
navigation / GUI parts of the class are not restated (out of scope).  Display-only work
(`np.copy` + `draw_border`, :929-934) is skipped -- it never feeds results.
"""
from __future__ import annotations

import numpy as np
import cv2

N_FEATURES = 700          # main.py:33,36
LOWE_RATIO = 0.7          # main.py:691
RANSAC_THRESH = 2.0       # main.py:857
BLUR_KSIZE = 31           # main.py:897-898


class RefMosaic:
    def __init__(self, first_image, output_height_times=2, output_width_times=1.2, detector_type="sift",
                 canvas_size=None, float64_canvas=True):
        # main.py:29-37 detector / matcher choice
        self.detector_type = detector_type
        if detector_type == "sift":
            self.detector = cv2.SIFT_create(N_FEATURES)
            self.bf = cv2.BFMatcher()
        elif detector_type == "orb":
            self.detector = cv2.ORB_create(N_FEATURES)
            self.bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
        else:
            raise ValueError(detector_type)
        # main.py:104-112
        self.frame_prev = first_image
        self.kp_prev, self.des_prev = self.detector.detectAndCompute(
            cv2.cvtColor(first_image, cv2.COLOR_BGR2GRAY), None)
        # main.py:80-81 canvas (float64 in the reference: np.zeros without dtype)
        fh, fw, fc = first_image.shape
        if canvas_size is None:
            ch, cw = int(output_height_times * fh), int(output_width_times * fw)
        else:
            ch, cw = canvas_size
        self.output_img = np.zeros((ch, cw, fc), dtype=np.float64 if float64_canvas else np.uint8)
        # main.py:86-90 (names are swapped in the reference: w_offset is the ROW offset)
        self.w_offset = int(ch / 1 - fh / 1)
        self.h_offset = int(cw / 2 - fw / 2)
        self.output_img[self.w_offset:self.w_offset + fh, self.h_offset:self.h_offset + fw, :] = first_image
        # main.py:92-94
        self.H_old = np.eye(3)
        self.H_old[0, 2] = self.h_offset
        self.H_old[1, 2] = self.w_offset
        # main.py:97-102
        self.stabilization_enabled = True
        self.homography_history = []
        self.history_size = 5
        self.translation_threshold = 50
        self.scale_threshold = 0.3
        self.last_status = "init"
        self.H = None

    # ---- main.py:676-698 ------------------------------------------------------------------
    def match(self, des_cur, des_prev):
        if self.detector_type == "sift":
            good = []
            for pair in self.bf.knnMatch(des_cur, des_prev, k=2):
                m, n = pair
                if m.distance < LOWE_RATIO * n.distance:
                    good.append(m)
        else:
            good = self.bf.match(des_cur, des_prev)
        return sorted(good, key=lambda d: d.distance)      # stable, main.py:698

    # ---- main.py:836-859 ------------------------------------------------------------------
    @staticmethod
    def findHomography(kp_a, kp_b, matches):
        pa = np.zeros((len(matches), 1, 2), dtype=np.float32)
        pb = np.zeros((len(matches), 1, 2), dtype=np.float32)
        for i, m in enumerate(matches):
            pa[i] = kp_a[m.queryIdx].pt
            pb[i] = kp_b[m.trainIdx].pt
        H, _mask = cv2.findHomography(pa, pb, cv2.RANSAC, ransacReprojThreshold=RANSAC_THRESH)
        return H

    # ---- main.py:761-801 (same decisions, same printed warnings) ------------------------------
    def validate_homography(self, H):
        if H is None:
            return False
        if np.any(np.isnan(H)) or np.any(np.isinf(H)):
            return False
        translation = np.sqrt(H[0, 2] ** 2 + H[1, 2] ** 2)
        with np.errstate(invalid="ignore"):
            scale = np.sqrt(np.linalg.det(H[:2, :2]))     # NaN when det<0 -> both tests False (quirk A.11)
        if translation > self.translation_threshold:
            print(f"Предупреждение: Обнаружено большое смещение ({translation:.1f}px), возможна тряска")          # :788
            return False
        if abs(scale - 1.0) > self.scale_threshold:
            print(f"Предупреждение: Обнаружено большое изменение масштаба ({scale:.2f}), возможна тряска")      # :793
            return False
        if abs(H[2, 0]) > 0.001 or abs(H[2, 1]) > 0.001:
            print("Предупреждение: Обнаружены сильные перспективные искажения")                                # :798
            return False
        return True

    # ---- main.py:803-834 ------------------------------------------------------------------
    def smooth_homography(self, H):
        if not self.stabilization_enabled:
            return H
        self.homography_history.append(H.copy())
        if len(self.homography_history) > self.history_size:
            self.homography_history.pop(0)
        if len(self.homography_history) < 2:
            return H
        w = np.linspace(0.5, 1.0, len(self.homography_history))
        w = w / np.sum(w)
        acc = np.zeros_like(H)
        for wi, h in zip(w, self.homography_history):
            acc += wi * h
        return acc

    # ---- main.py:861-927 ------------------------------------------------------------------
    def warp(self, frame_cur, H):
        ch, cw = self.output_img.shape[:2]
        warped = cv2.warpPerspective(frame_cur, H, (cw, ch), flags=cv2.INTER_LINEAR)
        self.output_img = blend_step_cv(self.output_img, warped)
        return self.output_img

    # ---- main.py:710-759 ------------------------------------------------------------------
    def process_frame(self, frame_cur, frame_count=0):
        self.frame_cur = frame_cur
        gray = cv2.cvtColor(frame_cur, cv2.COLOR_BGR2GRAY)
        self.kp_cur, self.des_cur = self.detector.detectAndCompute(gray, None)
        self.matches = self.match(self.des_cur, self.des_prev)
        if len(self.matches) < 4:
            print(f"Предупреждение: Недостаточно совпадений ({len(self.matches)}), пропуск кадра")          # :723
            self.last_status = "skip_few_matches"          # :722-724, state not advanced
            return
        H_rel = self.findHomography(self.kp_cur, self.kp_prev, self.matches)
        if H_rel is None:
            print("Предупреждение: Не удалось вычислить гомографию, пропуск кадра")                      # :730
            self.last_status = "skip_no_h"                 # :729-731
            return
        if not self.validate_homography(H_rel):
            print("Предупреждение: Невалидная гомография (тряска/размытие), использую последнюю валидную")  # :735
            H_rel = np.eye(3)                              # :734-737
            self.last_status = "rejected_identity"
        else:
            self.last_status = "ok"
        self.H_rel = H_rel
        H_s = self.smooth_homography(H_rel)                # :743
        self.H = np.matmul(self.H_old, H_s)                # :746
        self.warp(frame_cur, self.H)                       # :748
        self.H_old = self.H                                # :756-759
        self.kp_prev, self.des_prev, self.frame_prev = self.kp_cur, self.des_cur, self.frame_cur


def blend_step_cv(canvas, warped):
    """One blend step of `VideMosaic.warp` (main.py:878-927) on (canvas_before, warped) with cv2 / NumPy,
    dtype-for-dtype.  `canvas` may be float64 (reference) or uint8 (lossless: only integers 0..255 are stored).
    Returns the new canvas (same dtype as the input)."""
    mask_new = np.any(warped > 0, axis=2).astype(np.uint8) * 255
    mask_old = np.any(canvas > 0, axis=2).astype(np.uint8) * 255
    overlap = cv2.bitwise_and(mask_new, mask_old)
    out = canvas
    if np.any(overlap):
        dist_new = cv2.distanceTransform(mask_new, cv2.DIST_L2, 3)
        dist_old = cv2.distanceTransform(mask_old, cv2.DIST_L2, 3)
        dist_sum = dist_new + dist_old + 1e-6               # stays float32 (NEP 50)
        w_new = cv2.GaussianBlur((dist_new / dist_sum).astype(np.float32), (BLUR_KSIZE, BLUR_KSIZE), 0)
        w_old = cv2.GaussianBlur((dist_old / dist_sum).astype(np.float32), (BLUR_KSIZE, BLUR_KSIZE), 0)
        blended = (canvas.astype(np.float32) * w_old[:, :, None] + warped.astype(np.float32) * w_new[:, :, None])
        ov3 = (overlap > 0)[:, :, None]
        out = np.where(ov3, blended.astype(np.uint8), canvas)
        non_ov = (cv2.bitwise_and(mask_new, cv2.bitwise_not(overlap)) > 0)[:, :, None]
        out = np.where(non_ov, warped, out)
    else:
        out = canvas.copy()
        sel = warped > 0
        out[sel] = warped[sel]                              # channel-wise overwrite, main.py:927
    return out.astype(canvas.dtype, copy=False)
