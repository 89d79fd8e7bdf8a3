"""Golden fixture for the live-preview thumbnail.  /root/reference/gui.py cannot be imported in the build
container (customtkinter is not installed), so the fixture is produced by the call sequence of gui.py:143-158 itself
(oracle.preview.gui_thumbnail: astype(uint8) -> cv2.cvtColor(BGR2RGB) -> Image.fromarray -> resize((400, 300))) with the live
cv2 / Pillow of this image, on the final ORB canvas of the clip fixture that the unmodified reference produced.
Build container only:   python tests/golden/make_golden_preview.py"""
import hashlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

if __name__ == "__main__":
    import PIL
    from oracle.preview import gui_thumbnail
    canvas = np.load(HERE / "clip01_orb.npz")["canvas_final"]
    out = {"pillow_version": np.array(PIL.__version__)}
    thumb = gui_thumbnail(canvas)                               # 480 x 512 -> 300 x 400, both axes reduced
    out["thumb"] = thumb
    out["thumb_sha256"] = np.frombuffer(hashlib.sha256(thumb.tobytes()).digest(), np.uint8)
    small = np.ascontiguousarray(canvas[300:420, 180:340])      # 120 x 160 -> 300 x 400: enlargement (support 2, 5 taps)
    out["small_in"] = small
    out["small_thumb"] = gui_thumbnail(small)
    out["mixed_thumb"] = gui_thumbnail(canvas, size=(1100, 200))   # one axis enlarged, one reduced
    up = gui_thumbnail(canvas, size=(640, 600))                 # both axes enlarged; pinned by hash + a strided sample
    out["up_sha256"] = np.frombuffer(hashlib.sha256(up.tobytes()).digest(), np.uint8)
    out["up_sub4"] = up[::4, ::4].copy()
    np.savez_compressed(HERE / "preview.npz", **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})
