"""The launcher's class swap (b200mosaic.run) driven by a STAND-IN for the reference's `main` module that
follows main.main()'s sequence (main.py:1575-1670): cv2.VideoCapture loop -> process_frame -> crop_black_areas(output_img) ->
scale_to_screen -> cv2.imwrite('mosaic.jpg') -> output_img.astype(uint8).  /root/reference does not exist on the GPU box, so the
stand-in carries the reference's two finalisation functions as restated (and pinned) in oracle/finalize.py."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stand_in_main():
    import cv2
    from oracle import finalize as ofin
    ref = types.ModuleType("main")
    ref.cv2 = cv2
    ref.crop_black_areas = ofin.crop_black_areas
    ref.scale_to_screen = ofin.scale_to_screen
    ref.VideMosaic = None
    ref.log = {}

    def main(video_path, show_intermediate=False, output_dir=None, max_frames=40):
        cap = ref.cv2.VideoCapture(video_path)                               # main.py:1579 (the module global `cv2`, looked up at call time)
        ret, first = cap.read()
        vm = ref.VideMosaic(first, detector_type="sift", show_intermediate=show_intermediate, output_dir=output_dir)   # :1603
        n = 0
        while cap.isOpened() and n < max_frames:
            ret, frame = cap.read()                                          # :1597
            if not ret:
                break
            n += 1
            vm.process_frame(frame, n)                                       # :1613
        cap.release()
        cropped = ref.crop_black_areas(vm.output_img, threshold=80, margin=30)   # :1649
        ref.log["cropped_shape"] = cropped.shape
        scaled = ref.scale_to_screen(cropped)                                # :1656
        ref.cv2.imwrite(str(output_dir) + "/mosaic.jpg", scaled)             # :1663-1665
        ref.log["scaled"] = scaled
        ref.log["full"] = vm.output_img.astype(np.uint8)                     # :1670
        ref.log["vm"] = vm
        ref.log["frames"] = n
    ref.main = main
    return ref


@pytest.mark.parametrize("ahead", [True, False])
def test_launcher_swap_with_read_ahead_and_device_finalize(golden_dir, tmp_path, ahead):
    import cv2
    import b200mosaic
    from b200mosaic import run as brun
    from oracle import finalize as ofin
    ref = _stand_in_main()
    real_cap = cv2.VideoCapture
    brun.AheadCapture.current = None
    try:
        if ahead:
            brun.AheadCapture.real = real_cap
            cv2.VideoCapture = brun.AheadCapture
        ref.VideMosaic = brun.make_swapped_class("orb", ahead=ahead)
        brun.install_device_finalize(ref)
        brun.LazyCanvas.materialized = 0
        ref.main(str(golden_dir / "clip01.mp4"), output_dir=tmp_path)
    finally:
        cv2.VideoCapture = real_cap
    assert ref.log["frames"] == 40
    # mosaic.jpg was produced without a full-canvas D2H: the only materialisation is main()'s own .astype(uint8) for YOLO (:1670)
    assert brun.LazyCanvas.materialized == 1
    full = ref.log["full"]
    want = ofin.scale_to_screen(ofin.crop_black_areas(full, threshold=80, margin=30))
    assert np.array_equal(ref.log["scaled"], want)
    assert ref.log["cropped_shape"] == ofin.crop_black_areas(full, threshold=80, margin=30).shape
    # mosaic.jpg was encoded on the device (the launcher's cv2.imwrite proxy): the file cv2 itself would have written, byte for byte
    ok, enc = cv2.imencode(".jpg", ref.log["scaled"])
    assert ok and (tmp_path / "mosaic.jpg").read_bytes() == enc.tobytes()
    assert ref.cv2.imwrite(str(tmp_path / "other.png"), ref.log["scaled"]) and (tmp_path / "other.png").exists()   # everything else passes through
    # same mosaic with and without the reader thread, and equal to direct class use
    cap = real_cap(str(golden_dir / "clip01.mp4"))
    frames = [cap.read()[1] for _ in range(41)]
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    for t in range(1, 41):
        vm.process_frame(frames[t], t)
    assert np.array_equal(vm.output_img, full)
    assert np.array_equal(vm.H_old, ref.log["vm"].H_old)
