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
