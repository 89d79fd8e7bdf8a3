"""Debug helper (GPU box): dump the device SIFT output next to live cv2's for a few frames -> gpurun_out/sift_dump.npz"""
import sys
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200mosaic.ops as ops  # noqa: E402
from b200mosaic.synth import DroneSweep  # noqa: E402
from oracle import sift as osift  # noqa: E402

cv2.ipp.setUseIPP(False)
out = {}
cap = cv2.VideoCapture(str(ROOT / "tests/golden/clip01.mp4"))
frames = []
for i in range(8):
    frames.append(cap.read()[1])
cases = {"clip2": frames[2], "clip6": frames[6], "clip1": frames[1],
         "syn1080": DroneSweep(1920, 1080, seed=9, ground_size=2048).next(),
         "syn360": DroneSweep(640, 360, seed=9, ground_size=2048).next()}
for name, f in cases.items():
    g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    kp, des = ops.sift_detect_and_compute(torch.from_numpy(g).cuda())
    kc, dc = osift.cv_detect_and_compute(g)
    kall = cv2.SIFT_create(0).detect(g, None)
    out[name + "_gray"] = g
    out[name + "_kp"] = kp; out[name + "_des"] = des
    out[name + "_kc"] = kc; out[name + "_dc"] = dc
    out[name + "_kall"] = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kall])
    print(name, len(kp), len(kc), len(kall))
np.savez_compressed(ROOT / "gpurun_out" / "sift_dump.npz", **out)
