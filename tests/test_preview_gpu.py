"""bm_preview (device thumbnail of the live canvas) against the oracle and the committed fixture."""
import numpy as np
import pytest
from pathlib import Path

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _mosaic_with_canvas(canvas):
    """a handle whose canvas equals `canvas` exactly: frame 0 of the canvas size pasted over the whole canvas (main.py:104-110)"""
    import b200mosaic
    h, w = canvas.shape[:2]
    return b200mosaic.VideMosaic(canvas, detector_type="orb", show_intermediate=False, visualize=False, canvas_size=(h, w))


def test_preview_matches_fixture_and_oracle():
    import hashlib
    from oracle import preview as opv
    g = np.load(GOLD / "preview.npz")
    canvas = np.load(GOLD / "clip01_orb.npz")["canvas_final"]
    vm = _mosaic_with_canvas(canvas)
    assert np.array_equal(vm.output_img, canvas)
    got = vm.preview()
    assert got.shape == (300, 400, 3)
    assert np.array_equal(got, g["thumb"])                                             # both axes reduced
    assert np.array_equal(vm.preview((1100, 200)), g["mixed_thumb"])                   # one enlarged, one reduced
    up = vm.preview((640, 600))                                                        # both enlarged
    assert np.array_equal(np.frombuffer(hashlib.sha256(up.tobytes()).digest(), np.uint8), g["up_sha256"])
    assert np.array_equal(vm.preview((512, 480)), canvas[..., ::-1])                   # same size: Pillow copies, the filter is the identity
    assert np.array_equal(vm.preview((400, 300), rgb=False), opv.thumbnail(canvas, rgb=False))
    vm.close()


def test_preview_of_live_1080p_canvas():
    """after real stitching at BASELINE.json's size: thumbnail of the device canvas == oracle thumbnail of output_img"""
    import b200mosaic
    from b200mosaic.synth import DroneSweep
    from oracle import preview as opv
    frames = DroneSweep(1920, 1080, seed=11, ground_size=4096, max_step=10.0).frames(4)
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    for t in range(1, 4):
        vm.process_frame(frames[t], t)
        got = vm.preview()
        assert np.array_equal(got, opv.thumbnail(vm.output_img)), t
    assert np.array_equal(vm.preview((97, 61)), opv.thumbnail(vm.output_img, (97, 61)))
    vm.close()
