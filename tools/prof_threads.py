"""Is one stream host bound?  T host threads, each driving its own handle over the same device-resident sweep (ctypes releases the GIL
inside the library): aggregate frames/s for T = 1, 2, 4.   python tools/prof_threads.py [orb|sift] [frames]"""
import sys, time, threading
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200mosaic
from b200mosaic.synth import DroneSweep
det = sys.argv[1] if len(sys.argv) > 1 else "orb"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
w, h = 1920, 1080
frames = DroneSweep(w, h, seed=1234, ground_size=4096, max_step=12.0, max_travel=0.8 * h).frames(n + 4)
dev = torch.from_numpy(np.stack(frames)).cuda(); fb = h * w * 3
P = dev.data_ptr()
for T in (1, 2, 4):
    vms = [b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False) for _ in range(T)]
    for vm in vms:
        vm.warm_up()
        for i in range(1, 6):
            vm.process_frame_device(P + i * fb, P + (i + 1) * fb, P + (i + 2) * fb, P + (i + 3) * fb)
        vm.sync()
    bar = threading.Barrier(T + 1)
    def run(vm):
        bar.wait()
        for i in range(6, n + 1):
            vm.process_frame_device(P + i * fb, P + (i + 1) * fb, P + (i + 2) * fb, P + (i + 3) * fb)
        vm.sync()
        bar.wait()
    th = [threading.Thread(target=run, args=(vm,)) for vm in vms]
    for t in th: t.start()
    torch.cuda.synchronize(); bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    for t in th: t.join()
    print(det, "threads", T, "aggregate fps", T * (n - 5) / dt, "us/frame/stream", 1e6 * dt / (n - 5))
    for vm in vms: vm.close()
