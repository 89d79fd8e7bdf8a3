"""In-pipeline latencies (BM_PROFILE=1): python tools/prof_pipeline.py [orb|sift] [frames]"""
import os, sys, time
from pathlib import Path
os.environ.setdefault("BM_PROFILE", "1")   # BM_PROFILE=0 (exported before the call): no profiling events, they add host syncs
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200mosaic
from b200mosaic.synth import DroneSweep
det = sys.argv[1] if len(sys.argv) > 1 else "orb"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
w, h = 1920, 1080
frames = DroneSweep(w, h, seed=1234, ground_size=4096, max_step=12.0, max_travel=0.8 * h).frames(n + 5)
ahead = int(os.environ.get('BM_AHEAD', '3'))
dev = torch.from_numpy(np.stack(frames)).cuda(); fb = h * w * 3
vm = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
vm.warm_up()
for i in range(1, 6):
    vm.process_frame_device(dev.data_ptr() + i * fb, dev.data_ptr() + (i + 1) * fb, dev.data_ptr() + (i + 2) * fb if ahead > 1 else None, dev.data_ptr() + (i + 3) * fb if ahead > 2 else None)
vm.sync(); torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(6, n + 1):
    vm.process_frame_device(dev.data_ptr() + i * fb, dev.data_ptr() + (i + 1) * fb, dev.data_ptr() + (i + 2) * fb if ahead > 1 else None, dev.data_ptr() + (i + 3) * fb if ahead > 2 else None)
vm.sync(); dt = time.perf_counter() - t0
print(det, "fps", (n - 5) / dt, "us/frame", 1e6 * dt / (n - 5))
vm.close()
