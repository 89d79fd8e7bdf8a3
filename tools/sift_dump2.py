"""Debug helper (GPU box): ALL SIFT keypoints (no retainBest) ours vs cv2 for a few clip frames -> gpurun_out/sift_all.npz"""
import sys
from pathlib import Path
import cv2, numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200mosaic.ops as ops
cv2.ipp.setUseIPP(False)
cap = cv2.VideoCapture(str(ROOT / "tests/golden/clip01.mp4"))
out = {}
for t in range(120):
    ok, f = cap.read()
    if t not in (0, 53, 106):
        continue
    g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    kp, des = ops.sift_detect_and_compute(torch.from_numpy(g).cuda(), 100000)
    k7, d7 = ops.sift_detect_and_compute(torch.from_numpy(g).cuda(), 700)
    kall = cv2.SIFT_create(0).detect(g, None)
    k700 = cv2.SIFT_create(700).detect(g, None)
    conv = lambda ks: np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in ks])
    out[f"ours_all_{t}"] = kp; out[f"ours_700_{t}"] = k7; out[f"cv_all_{t}"] = conv(kall); out[f"cv_700_{t}"] = conv(k700)
    print(t, len(kp), len(kall), len(k7), len(k700))
np.savez_compressed(ROOT / "gpurun_out" / "sift_all.npz", **out)
