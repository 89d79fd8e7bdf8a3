"""Profiling helper: a few frames of the per-frame path for one detector (run under ncu / compute-sanitizer).
    python tools/prof_frames.py [orb|sift] [frames] [WxH]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200mosaic  # noqa: E402
from b200mosaic.synth import DroneSweep  # noqa: E402

det = sys.argv[1] if len(sys.argv) > 1 else "orb"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w, h = map(int, (sys.argv[3] if len(sys.argv) > 3 else "1920x1080").split("x"))
frames = DroneSweep(w, h, seed=1234, ground_size=4096, max_step=12.0, max_travel=0.8 * h).frames(n + 1)
dev = torch.from_numpy(np.stack(frames)).cuda()
fb = h * w * 3
vm = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
vm.set_overlap(False)
for i in range(1, n + 1):
    st = vm.process_frame_device(dev.data_ptr() + i * fb)
    assert st == 0, st
vm.sync()
print("ok", det, n, "frames")
