"""GPU parity of the mosaic finalisation: bm_finalize == crop_black_areas + scale_to_screen of the
reference (main.py:980-1038, 1647-1659) on the same canvas, bit for bit (oracle/finalize.py, pinned by tests/golden/finalize.npz)."""
import numpy as np
import pytest

from oracle import finalize as fin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mosaic(golden_dir):
    import b200mosaic
    frames = np.load(golden_dir / "clip01_frames.npz")["frames"]
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    for t in range(1, len(frames)):
        vm.process_frame(frames[t], t)
    return vm


@pytest.mark.parametrize("thr,margin,target", [(80, 30, None), (15, 5, None), (80, 30, (320, 300)), (0, 0, (1000, 100)), (40, 7, (3000, 3000))])
def test_finalize_matches_reference_functions(mosaic, thr, margin, target):
    canvas = mosaic.output_img
    tw, th = target if target else (None, None)
    want = fin.scale_to_screen(fin.crop_black_areas(canvas, threshold=thr, margin=margin), tw, th)      # the reference's own calls
    got = mosaic.finalize(thr, margin, tw, th)
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want)
    x, y, w, h = mosaic.last_crop_rect
    assert (x, y, w, h) == fin.crop_rect(canvas, thr, margin)


def test_finalize_whole_canvas_and_exact_half(mosaic):
    canvas = mosaic.output_img
    hc, wc = canvas.shape[:2]
    # nothing above the threshold -> crop_black_areas returns the image itself (main.py:996-997)
    got = mosaic.finalize(255, 30, 1234, 777)
    assert mosaic.last_crop_rect == (0, 0, wc, hc)
    assert np.array_equal(got, fin.scale_to_screen(canvas, 1234, 777))
    # exact 2x2 decimation: cv2 routes INTER_LINEAR to INTER_AREA
    got = mosaic.finalize(255, 0, wc // 2, hc // 2)
    assert got.shape == (hc // 2, wc // 2, 3)
    assert np.array_equal(got, fin.scale_to_screen(canvas, wc // 2, hc // 2))


def test_finalize_empty_crop_is_an_error(mosaic):
    import b200mosaic
    with pytest.raises(b200mosaic.B200MosaicError):
        mosaic.finalize(80, 5000)
