"""Golden fixture for the mosaic finalisation: runs the UNMODIFIED reference functions
`crop_black_areas` / `scale_to_screen` of /root/reference/main.py on the final ORB canvas of the clip fixture (and on an
upscaling case).  Build container only:   python tests/golden/make_golden_finalize.py"""
import hashlib
from pathlib import Path

import numpy as np

from make_golden import load_reference

HERE = Path(__file__).resolve().parent

if __name__ == "__main__":
    ref = load_reference()
    canvas = np.load(HERE / "clip01_orb.npz")["canvas_final"]
    out = {}
    for name, (thr, margin) in {"main": (80, 30), "default": (15, 5)}.items():
        crop = ref.crop_black_areas(canvas, threshold=thr, margin=margin)
        scaled = ref.scale_to_screen(crop)
        out[f"{name}_crop_shape"] = np.array(crop.shape)
        out[f"{name}_crop_first"] = crop[0, 0].copy()
        out[f"{name}_scaled_shape"] = np.array(scaled.shape)
        out[f"{name}_scaled_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(scaled).tobytes()).digest(), np.uint8)
        out[f"{name}_scaled_sub8"] = scaled[::8, ::8].copy()      # strided sample (the full image is pinned by its hash)
    small = ref.scale_to_screen(canvas, target_w=320, target_h=300)
    out["small_scaled"] = small
    np.savez_compressed(HERE / "finalize.npz", **out)
    print({k: v.shape for k, v in out.items()})
