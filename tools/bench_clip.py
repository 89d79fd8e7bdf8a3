#!/usr/bin/env python
"""BASELINE.json configs 1 / 2 end to end WITH decode (SURVEY.md 8d "Config 1 / 2", 8f rank 2): the reference's own clip
(tests/golden/clip01.mp4: 854x480, 592 frames; the survey measured 4.49 fps SIFT / 6.58 fps ORB for the reference on 8 cores) through
the driver loop of main.main() (main.py:1575-1666):

    cap = cv2.VideoCapture(path); ret, first = cap.read(); vm = VideMosaic(first, detector_type=...)
    while cap.isOpened(): ret, frame = cap.read(); vm.process_frame(frame, n)
    cropped = crop_black_areas(vm.output_img, 80, 30); scaled = scale_to_screen(cropped); cv2.imwrite('mosaic.jpg', scaled)

/root/reference does not exist on the GPU box, so the loop is restated here (the same stand-in tests/test_run_gpu.py drives); the
class swap, the reader thread (run.AheadCapture), the device finalisation and the device JPEG encoder are the launcher's own
(b200mosaic.run).  Timed per detector, wall clock, everything inside: container open, H.264 decode, H2D, all kernels, per-frame result
read-back, finalisation, mosaic.jpg on disk.  Beside it: the decode alone (the floor of any implementation that keeps cv2.VideoCapture)
and the CPU port (oracle.mosaic_ref.RefMosaic + the restated finalisation) over the first --cpu-frames frames of the same loop.

    python tools/bench_clip.py [--cpu-frames 40] [--repeat 3] [--out gpurun_out/clip.json]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
CLIP = ROOT / "tests" / "golden" / "clip01.mp4"


def driver_module(cv2, crop, scale):
    """a module object with main.main()'s sequence (main.py:1575-1666); `cv2`, `VideMosaic`, `crop_black_areas`, `scale_to_screen` are
    module globals looked up at call time exactly like in the reference, so the launcher's swaps apply"""
    ref = types.ModuleType("main")
    ref.cv2, ref.crop_black_areas, ref.scale_to_screen, ref.VideMosaic = cv2, crop, scale, None
    ref.log = {}

    def main(video_path, detector, output_dir, max_frames=None):
        tm = time.perf_counter
        t_a = tm()
        cap = ref.cv2.VideoCapture(video_path)                                   # main.py:1579
        ret, first = cap.read()
        vm = ref.VideMosaic(first, detector_type=detector, show_intermediate=False, output_dir=output_dir, visualize=False)   # :1603
        n = 0
        t_b = tm()
        t_read = t_proc = 0.0
        while cap.isOpened() and (max_frames is None or n < max_frames):
            t0 = tm()
            ret, frame = cap.read()                                              # :1597
            t1 = tm()
            t_read += t1 - t0
            if not ret:
                break
            n += 1
            vm.process_frame(frame, n)                                           # :1613
            t_proc += tm() - t1
        cap.release()
        t_c = tm()
        cropped = ref.crop_black_areas(vm.output_img, threshold=80, margin=30)   # :1649
        scaled = ref.scale_to_screen(cropped)                                    # :1656
        ref.cv2.imwrite(os.path.join(output_dir, "mosaic.jpg"), scaled)          # :1663-1665
        ref.log.update(frames=n, vm=vm, scaled=scaled,
                       split={"setup_s": t_b - t_a, "read_s": t_read, "process_frame_s": t_proc, "finalize_and_write_s": tm() - t_c})
    ref.main = main
    return ref


def decode_only(cv2, max_frames=None):
    t0 = time.perf_counter()
    cap = cv2.VideoCapture(str(CLIP))
    n = 0
    while max_frames is None or n <= max_frames:
        ok, _ = cap.read()
        if not ok:
            break
        n += 1
    cap.release()
    return n - 1, time.perf_counter() - t0


def run_b200(det, repeat, ahead=True):
    """the launcher's configuration (b200mosaic.run.main) around the stand-in driver; returns the per-run records"""
    import ctypes as C
    import cv2
    from b200mosaic import run as brun, _lib
    lib = _lib.load()

    def not_on_device(*a, **k):                   # the driver's own host crop / scale: reaching them means the device path was not taken
        raise RuntimeError("finalisation fell back to the host functions")
    runs = []
    for _ in range(repeat):
        ref = driver_module(cv2, not_on_device, not_on_device)
        real_cap = cv2.VideoCapture
        brun.AheadCapture.current = None
        with tempfile.TemporaryDirectory() as td:
            try:
                if ahead:
                    brun.AheadCapture.real = real_cap
                    cv2.VideoCapture = brun.AheadCapture
                ref.VideMosaic = brun.make_swapped_class(det, ahead=ahead)
                brun.install_device_finalize(ref)
                brun.LazyCanvas.materialized = 0
                lm = np.zeros(3, np.uint64)
                lib.bm_debug_lm_stats(lm.ctypes.data_as(C.c_void_p), 1)
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()) as warn:
                    ref.main(str(CLIP), det, td)
                dt = time.perf_counter() - t0
            finally:
                cv2.VideoCapture = real_cap
            jpg = (Path(td) / "mosaic.jpg").read_bytes()
        vm = ref.log.pop("vm")
        lib.bm_debug_lm_stats(lm.ctypes.data_as(C.c_void_p), 0)
        ok, enc = cv2.imencode(".jpg", ref.log["scaled"])
        runs.append({"frames": ref.log["frames"], "seconds": dt, "fps": ref.log["frames"] / dt, "mosaic_jpg_bytes": len(jpg),
                     "mosaic_jpg_identical_to_cv2_imencode": bool(ok and enc.tobytes() == jpg),
                     "full_canvas_d2h": brun.LazyCanvas.materialized, "warnings_printed": warn.getvalue().count("\n"),
                     "split": ref.log["split"], "polish": {"runs": int(lm[0]), "lm_iterations": int(lm[1]), "by_eigen_decomposition": int(lm[2])}})
        vm.close()
    return runs


def run_cpu(det, max_frames, cores):
    """the CPU port through the same driver loop (decode inside), first `max_frames` frames"""
    import cv2
    from oracle import finalize as ofin
    from oracle.mosaic_ref import RefMosaic
    ipp0, thr0 = cv2.ipp.useIPP(), cv2.getNumThreads()
    cv2.setNumThreads(cores)
    cv2.ipp.setUseIPP(True)                       # timing runs keep IPP on; the caller's setting is restored below
    ref = driver_module(cv2, ofin.crop_black_areas, ofin.scale_to_screen)
    ref.VideMosaic = lambda first, detector_type, **kw: RefMosaic(first, detector_type=detector_type)
    try:
        with tempfile.TemporaryDirectory() as td, contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            ref.main(str(CLIP), det, td, max_frames=max_frames)
            dt = time.perf_counter() - t0
    finally:
        cv2.ipp.setUseIPP(ipp0)
        cv2.setNumThreads(thr0)
    return {"frames": ref.log["frames"], "seconds": dt, "fps": ref.log["frames"] / dt, "cores": cores, "kind": "port",
            "sample": f"first {ref.log['frames']} frames of the clip through the same loop (decode, finalisation and mosaic.jpg inside), "
                      f"oracle.mosaic_ref.RefMosaic, cv2 {cv2.__version__}, cv2.setNumThreads({cores}), IPP on"}


def clip_record(cpu_frames=40, repeat=3, detectors=("sift", "orb")):
    import cv2
    rec = {"what": "BASELINE configs 1 / 2: tests/golden/clip01.mp4 (854x480, 592 frames, default canvas 960x1024) through main.main()'s loop "
                   "(main.py:1575-1666) with the launcher's swaps: cv2.VideoCapture decode on the reader thread, H2D, all kernels, per-frame "
                   "read-back, device finalisation, mosaic.jpg written; wall clock around the whole run incl. handle creation",
           "survey_reference_fps_8_cores": {"sift": 4.49, "orb": 6.58}}
    nd, td = decode_only(cv2)
    nd, td = decode_only(cv2)                   # second pass: file cached
    rec["decode_only"] = {"frames": nd, "seconds": td, "fps": nd / td, "what": "cv2.VideoCapture.read() of every frame, nothing else"}
    cores = os.cpu_count() or 1
    for det in detectors:
        runs = run_b200(det, repeat + 1)[1:]     # the first run pays module import / context / graph capture of a cold process
        runs.sort(key=lambda r: r["fps"])
        med = runs[len(runs) // 2]
        rec[det] = {"fps": med["fps"], "seconds": med["seconds"], "frames": med["frames"], "fps_min_max": [runs[0]["fps"], runs[-1]["fps"]],
                    "mosaic_jpg_bytes": med["mosaic_jpg_bytes"],
                    "mosaic_jpg_identical_to_cv2_imencode": all(r["mosaic_jpg_identical_to_cv2_imencode"] for r in runs),
                    "full_canvas_d2h": med["full_canvas_d2h"], "fraction_of_decode_rate": med["fps"] / (nd / td),
                    "polish": med["polish"], "split": med["split"],
                    "runs": [{"fps": r["fps"], "seconds": r["seconds"], **r["split"]} for r in runs]}
    for det in detectors:                        # after every device run: cv2's 16-thread pool keeps spinning for a while and would steal the
        if cpu_frames > 0:                       # cores the decode thread and the launch thread of the next device run need
            rec[det]["cpu_baseline"] = run_cpu(det, cpu_frames, cores)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-frames", type=int, default=40)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rec = clip_record(a.cpu_frames, a.repeat)
    s = json.dumps(rec)
    print(s, flush=True)
    if a.out:
        Path(a.out).write_text(s + "\n")


if __name__ == "__main__":
    main()
