"""Host pieces of the configs 1 / 2 bench leg (tools/bench_clip.py): the restated driver loop of main.main() (main.py:1575-1666) run on
the CPU port over the first frames of the reference's clip -- decode, stitching, finalisation and mosaic.jpg inside the timed run."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))


def test_driver_loop_on_the_cpu_port():
    import cv2
    import bench_clip
    n, dt = bench_clip.decode_only(cv2, 5)
    assert n == 5 and dt > 0
    rec = bench_clip.run_cpu("orb", 3, min(4, os.cpu_count() or 1))
    assert rec["frames"] == 3 and rec["fps"] > 0 and rec["kind"] == "port"


def test_device_leg_refuses_host_finalisation():
    """the device leg hands the driver crop / scale functions that raise: if the launcher's device finalisation were not taken, the run
    would fail loudly instead of timing the host path"""
    import inspect
    import bench_clip
    src = inspect.getsource(bench_clip.run_b200)
    assert "not_on_device" in src and "RuntimeError" in src


def test_clip_record_assembly(monkeypatch):
    """the record bench.py embeds (`clip`) is assembled from per-run dictionaries: median run, min / max, per-run time split, CPU legs
    after every device run -- checked here with stand-in runs (the device leg itself needs a GPU: tests/test_run_gpu.py)"""
    import json
    import bench_clip
    calls = []

    def fake_b200(det, repeat, ahead=True):
        calls.append(("gpu", det))
        return [{"frames": 591, "seconds": 0.4 + 0.1 * i, "fps": 591 / (0.4 + 0.1 * i), "mosaic_jpg_bytes": 1000, "mosaic_jpg_identical_to_cv2_imencode": True,
                 "full_canvas_d2h": 0, "warnings_printed": 0, "split": {"setup_s": 0.05, "read_s": 0.0, "process_frame_s": 0.3, "finalize_and_write_s": 0.01},
                 "polish": {"runs": 591, "lm_iterations": 2500, "by_eigen_decomposition": 0}} for i in range(repeat)]

    def fake_cpu(det, max_frames, cores):
        calls.append(("cpu", det))
        return {"frames": max_frames, "seconds": 1.0, "fps": float(max_frames), "cores": cores, "kind": "port", "sample": "stand-in"}
    monkeypatch.setattr(bench_clip, "run_b200", fake_b200)
    monkeypatch.setattr(bench_clip, "run_cpu", fake_cpu)
    rec = bench_clip.clip_record(cpu_frames=5, repeat=3)
    json.dumps(rec)
    assert [c[0] for c in calls] == ["gpu", "gpu", "cpu", "cpu"]                      # every device run before any CPU leg
    for det in ("sift", "orb"):
        r = rec[det]
        assert len(r["runs"]) == 3 and r["fps_min_max"][0] <= r["fps"] <= r["fps_min_max"][1]
        assert r["runs"][0]["process_frame_s"] == 0.3 and r["polish"]["by_eigen_decomposition"] == 0
        assert r["cpu_baseline"]["frames"] == 5
