"""Launcher: run the UNMODIFIED reference driver (`main.main`, main.py:1512-1717) with `VideMosaic` swapped for the
B200 implementation.

    python -m b200mosaic.run <video> [--reference-dir DIR] [--output-dir D] [--detector sift|orb]

`--reference-dir` is the checkout that holds the reference's main.py (it is imported, never modified).  YOLO detection,
A* navigation, cropping / scaling / mosaic.jpg writing all stay on the reference's own code path."""
from __future__ import annotations

import argparse
import importlib.util
import sys
import types
from pathlib import Path


def _stub_missing(name, attrs=()):
    try:
        __import__(name)
    except Exception:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            n = ".".join(parts[:i])
            sys.modules.setdefault(n, types.ModuleType(n))
        for a in attrs:
            setattr(sys.modules[name], a, object)


def load_reference_main(ref_dir: Path):
    # out-of-scope imports of main.py that may be absent on a GPU box; both are only used by detection / navigation
    try:
        import ultralytics  # noqa: F401
    except Exception:
        u = types.ModuleType("ultralytics")

        class _YOLO:
            def __init__(self, *a, **k):
                raise RuntimeError("ultralytics unavailable")      # caught at main.py:45-47, 68-70
        u.YOLO = _YOLO
        sys.modules["ultralytics"] = u
    _stub_missing("pathfinding.core.grid", ("Grid",))
    _stub_missing("pathfinding.core.diagonal_movement", ("DiagonalMovement",))
    _stub_missing("pathfinding.finder.a_star", ("AStarFinder",))
    spec = importlib.util.spec_from_file_location("main", str(ref_dir / "main.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["main"] = mod
    spec.loader.exec_module(mod)
    return mod


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("video")
    ap.add_argument("--reference-dir", default=".")
    ap.add_argument("--output-dir", default=None)
    ap.add_argument("--detector", default="sift", choices=["sift", "orb"])
    a = ap.parse_args(argv)
    import cv2
    import b200mosaic
    ref = load_reference_main(Path(a.reference_dir))
    try:                                    # headless OpenCV wheels raise in highgui calls (main.py:1616 calls waitKey)
        cv2.waitKey(1)
    except cv2.error:
        cv2.waitKey = lambda *x, **k: -1
        cv2.destroyAllWindows = cv2.imshow = cv2.namedWindow = lambda *x, **k: None
    b200mosaic.VideMosaic.reference_class = ref.VideMosaic
    det = a.detector

    class _Swapped(b200mosaic.VideMosaic):
        def __init__(self, first_image, *args, **kw):
            kw["detector_type"] = det       # main() hard-codes "sift" (main.py:1603)
            kw.setdefault("visualize", False)
            super().__init__(first_image, *args, **kw)

    ref.VideMosaic = _Swapped
    ref.main(video_path=a.video, show_intermediate=False, output_dir=a.output_dir)


if __name__ == "__main__":
    main()
