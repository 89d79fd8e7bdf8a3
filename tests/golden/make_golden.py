"""Generates the golden fixtures in this directory by running the UNMODIFIED reference class
(/root/reference/main.py `VideMosaic`) in the build container.  The reference cannot travel to the GPU box, the
fixtures can.  Usage (build container only):   python tests/golden/make_golden.py

Recipe: stub the two missing out-of-scope imports (ultralytics, pathfinding), import main.py by
path, construct with show_intermediate=False / visualize=False (headless cv2).  cv2.ipp.setUseIPP(False) so that
distanceTransform is the exactly specifiable integer chamfer.
Inputs: frames 0,4,8,12,16 of Data/'поиски квадрокоптера 2 (360p) 01.mp4' (854x480), downscaled to 427x240 with
INTER_AREA to keep the fixtures small.
"""
import glob
import sys
import types
import importlib.util
from pathlib import Path

import numpy as np
import cv2

HERE = Path(__file__).resolve().parent


def load_reference():
    u = types.ModuleType("ultralytics")

    class _YOLO:
        def __init__(self, *a, **k):
            raise RuntimeError("ultralytics unavailable (stub)")
    u.YOLO = _YOLO
    sys.modules["ultralytics"] = u
    for n in ("pathfinding", "pathfinding.core", "pathfinding.core.grid", "pathfinding.core.diagonal_movement",
              "pathfinding.finder", "pathfinding.finder.a_star"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["pathfinding.core.grid"].Grid = object
    sys.modules["pathfinding.core.diagonal_movement"].DiagonalMovement = object
    sys.modules["pathfinding.finder.a_star"].AStarFinder = object
    spec = importlib.util.spec_from_file_location("ref_main", "/root/reference/main.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def read_frames(idx, size=(427, 240)):
    path = [p for p in glob.glob("/root/reference/Data/*.mp4") if "(360p) 01" in p][0]
    cap = cv2.VideoCapture(path)
    out, i = [], 0
    while True:
        ok, f = cap.read()
        if not ok:
            break
        if i in idx:
            out.append(cv2.resize(f, size, interpolation=cv2.INTER_AREA))
        i += 1
        if i > max(idx):
            break
    return out


def kp_array(kps):
    return np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], dtype=np.float64)


def main():
    cv2.ipp.setUseIPP(False)
    ref = load_reference()
    frames = read_frames({0, 4, 8, 12, 16})
    np.savez_compressed(HERE / "clip01_frames.npz", frames=np.stack(frames))
    for det in ("orb", "sift"):
        vm = ref.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
        rec = {"kp0": kp_array(vm.kp_prev), "des0": vm.des_prev,
               "canvas0": vm.output_img.astype(np.uint8), "offsets": np.array([vm.h_offset, vm.w_offset])}
        Hs, Hrel, nm = [], [], []
        for t, f in enumerate(frames[1:], 1):
            canvas_before = vm.output_img.astype(np.uint8)
            vm.process_frame(f, t)
            Hs.append(vm.H.copy())
            nm.append(len(vm.matches))
            if t == 1:
                rec["kp1"] = kp_array(vm.kp_cur)
                rec["des1"] = vm.des_cur
                rec["matches1"] = np.array([[m.queryIdx, m.trainIdx, m.distance] for m in vm.matches], dtype=np.float64)
                rec["canvas_before1"] = canvas_before
                rec["canvas_after1"] = vm.output_img.astype(np.uint8)
        assert np.array_equal(vm.output_img, np.floor(vm.output_img))       # float64 canvas holds integers only
        rec["H"] = np.stack(Hs)
        rec["n_matches"] = np.array(nm)
        rec["history"] = np.stack(vm.homography_history)
        rec["canvas_final"] = vm.output_img.astype(np.uint8)
        np.savez_compressed(HERE / f"clip01_{det}.npz", **rec)
        print(det, "matches", nm, "kp", len(vm.kp_cur))


if __name__ == "__main__":
    main()
