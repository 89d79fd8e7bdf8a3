"""Profiling helper: the cv2-order retainBest emulation alone on a few list sizes (run under ncu)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from b200mosaic import ops
rng = np.random.default_rng(0)
for n, k, u8 in [(80000, 304, True), (40000, 304, True), (10000, 304, True), (2048, 304, True), (1000, 304, True), (400, 152, False), (62000, 700, False)]:
    r = rng.integers(20, 255, n).astype(np.float32) if u8 else rng.random(n).astype(np.float32)
    m = ops.cv_retain_best(r, k, as_u8=u8)
    print(n, k, u8, len(m))
