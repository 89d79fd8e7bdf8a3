#!/usr/bin/env python
"""CPU only: the restated cv2.findHomography (oracle/ransac.py, nine-parameter LM polish) against cv2 itself on EVERY consensus problem of the
reference's clip -- the CPU port (oracle.mosaic_ref.RefMosaic) is run over tests/golden/clip01.mp4 for ORB and SIFT, each frame's matched point
sets are intercepted at findHomography and solved both ways.  Prints the corner reprojection distance (854 x 480 frame) per detector.

    python tools/validate_polish_on_clip.py            (about 4 minutes on 8 cores)

Result of the run behind DESIGN.md section 2.1: ORB 591 frames, max 3.4e-6 px, median 9.5e-10 px (frame 359, the ill-conditioned one: 1.4e-8 px);
SIFT 591 frames, max 3.5e-7 px, median 3e-13 px.  (With np.linalg.eigh in place of the restated cv2 Jacobi: ORB max 8.9e-4 px at frame 359.)"""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import cv2

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ransac as orc            # noqa: E402
from oracle.mosaic_ref import RefMosaic     # noqa: E402

CORNERS = np.array([[0, 0, 1], [853, 0, 1], [853, 479, 1], [0, 479, 1.0]]).T


def reproj(Ha, Hb):
    a = Ha @ CORNERS; b = Hb @ CORNERS
    return float(np.abs(a[:2] / a[2] - b[:2] / b[2]).max())


def main():
    cv2.ipp.setUseIPP(False)
    for det in ("orb", "sift"):
        cap = cv2.VideoCapture(str(ROOT / "tests" / "golden" / "clip01.mp4"))
        ok, f0 = cap.read()
        m = RefMosaic(f0, detector_type=det)
        m.warp = lambda frame, H: m.output_img          # the trajectory only: the canvas does not feed back into the estimates
        errs = []
        inner = RefMosaic.findHomography

        def find(kp_a, kp_b, matches):
            src = np.float32([kp_a[x.queryIdx].pt for x in matches]); dst = np.float32([kp_b[x.trainIdx].pt for x in matches])
            Hcv = inner(kp_a, kp_b, matches)
            Ho = orc.find_homography_ransac(src, dst, 2.0)
            if Hcv is not None and Ho is not None:
                errs.append(reproj(Ho, Hcv))
            else:
                assert Hcv is None and Ho is None
            return Hcv
        m.findHomography = find
        n = 0
        with contextlib.redirect_stdout(io.StringIO()):
            while True:
                ok, f = cap.read()
                if not ok:
                    break
                n += 1
                m.process_frame(f, n)
        e = np.array(errs)
        print(f"{det}: {len(e)} consensus problems, restatement vs cv2 at the frame corners: max {e.max():.2e} px (frame {int(e.argmax()) + 1}), "
              f"median {np.median(e):.2e} px, {(e > 1e-3).sum()} above 1e-3 px")


if __name__ == "__main__":
    main()
