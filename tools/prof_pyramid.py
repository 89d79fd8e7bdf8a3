"""SIFT pyramid alone (bm_sift_pyramid_ms): python tools/prof_pyramid.py [WxH] [reps]   (BM_SIFT_TMA=1 selects the TMA form of the blur kernels)"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from b200mosaic import _lib
from b200mosaic.synth import DroneSweep
w, h = map(int, (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = _lib.load()
f = DroneSweep(w, h, seed=1234, ground_size=4096).next()
g = torch.from_numpy(np.ascontiguousarray(f[:, :, 1])).cuda()
ms = C.c_double(0); by = C.c_double(0)
st = lib.bm_sift_pyramid_ms(C.c_void_p(g.data_ptr()), h, w, reps, C.byref(ms), C.byref(by))
if st != 0:
    print("error:", lib.bm_last_error().decode()); sys.exit(1)
print("tma" if os.environ.get("BM_SIFT_TMA") else "cp.async", w, h, "status", st, "pyramid us/frame", 1e3 * ms.value, "GB/s", by.value / 1e9 / (ms.value / 1e3))
