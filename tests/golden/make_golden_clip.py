"""Full-clip goldens (BASELINE configs 1/2): runs the UNMODIFIED reference class (/root/reference/main.py `VideMosaic`) over ALL
frames of Data/'поиски квадрокоптера 2 (360p) 01.mp4' (854x480, the clip SURVEY 8d substitutes for the missing '03') for
detector_type = "orb" and "sift", in the build container.  Usage (build container only):

    python tests/golden/make_golden_clip.py

Writes next to this file
  clip01.mp4                  the clip itself (a reference-held input vector; test fixture only)
  clip01_full_<det>.npz       per frame: status (0 ok, 1 <4 matches, 2 no H, 3 rejected -> identity; classified from the
                              reference's own printed warnings, main.py:723,730,735), n_kp, n_matches, absolute H (main.py:746),
                              raw relative H when it was accepted (last_valid_H, main.py:740); the canvas at every 100th frame
                              (2x INTER_AREA downscale) and the final canvas at full resolution; CRC32 of every decoded frame so
                              that a test can tell a different video decode from a different result.
cv2.ipp.setUseIPP(False): the exactly specifiable integer chamfer distanceTransform.
"""
import contextlib
import glob
import io
import shutil
import sys
import zlib
from pathlib import Path

import numpy as np
import cv2

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import load_reference  # noqa: E402


def classify(text: str) -> int:
    if "Недостаточно совпадений" in text:
        return 1
    if "Не удалось вычислить гомографию" in text:
        return 2
    if "Невалидная гомография" in text:
        return 3
    return 0


def main():
    cv2.ipp.setUseIPP(False)
    ref = load_reference()
    src = [p for p in glob.glob("/root/reference/Data/*.mp4") if "(360p) 01" in p][0]
    shutil.copyfile(src, HERE / "clip01.mp4")
    cap = cv2.VideoCapture(str(HERE / "clip01.mp4"))
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    crc = np.array([zlib.crc32(f.tobytes()) for f in frames], dtype=np.uint32)
    print(len(frames), "frames", frames[0].shape)
    for det in ("orb", "sift"):
        vm = ref.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
        n = len(frames)
        status = np.zeros(n, np.int32); nkp = np.zeros(n, np.int32); nm = np.zeros(n, np.int32)
        H = np.zeros((n, 3, 3)); Hrel = np.full((n, 3, 3), np.nan)
        H[0] = vm.H_old; nkp[0] = len(vm.kp_prev)
        ckpt_idx, ckpt = [], []
        for t in range(1, n):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                vm.process_frame(frames[t], t)
            status[t] = classify(buf.getvalue())
            nkp[t] = len(vm.kp_cur); nm[t] = len(vm.matches)
            H[t] = vm.H_old
            if status[t] == 0:
                Hrel[t] = vm.last_valid_H
            if t % 100 == 0:
                ckpt_idx.append(t)
                c = vm.output_img.astype(np.uint8)
                ckpt.append(cv2.resize(c, (c.shape[1] // 2, c.shape[0] // 2), interpolation=cv2.INTER_AREA))
        np.savez_compressed(HERE / f"clip01_full_{det}.npz", status=status, n_kp=nkp, n_matches=nm, H=H, H_rel=Hrel,
                            ckpt_idx=np.array(ckpt_idx), ckpt=np.stack(ckpt), canvas_final=vm.output_img.astype(np.uint8),
                            frame_crc=crc, cv2_version=np.array(cv2.__version__))
        print(det, "status counts", np.bincount(status, minlength=4), "matches min/median", nm[1:].min(), int(np.median(nm[1:])))


if __name__ == "__main__":
    main()
