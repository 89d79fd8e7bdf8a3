"""The launcher's hooks around the reference's finalisation (b200mosaic.run.install_device_finalize / install_device_imwrite) have no
silent host fallback: what they take over runs on the device or raises; what they do not take over is passed to the reference untouched.
Checked here WITHOUT a CUDA device, where every device call must fail loudly."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour without a CUDA device")


def _ref_module():
    import cv2
    ref = types.ModuleType("main")
    ref.cv2 = cv2
    ref.calls = []
    ref.crop_black_areas = lambda image, threshold=15, margin=5: (ref.calls.append("crop"), image)[1]
    ref.scale_to_screen = lambda image, target_w=None, target_h=None: (ref.calls.append("scale"), image)[1]
    return ref


def test_imwrite_proxy_passes_through_or_raises(tmp_path):
    from b200mosaic import run as brun
    ref = _ref_module()
    brun.install_device_imwrite(ref)
    img = np.full((32, 48, 3), 128, np.uint8)
    assert ref.cv2.imwrite(str(tmp_path / "a.png"), img) and (tmp_path / "a.png").exists()            # other formats: the real cv2.imwrite
    assert ref.cv2.imwrite(str(tmp_path / "b.jpg"), img, [1, 90]) and (tmp_path / "b.jpg").exists()   # explicit parameters: the real one
    assert ref.cv2.imwrite(str(tmp_path / "g.jpg"), img[:, :, 0])                                       # not a 3-channel image: the real one
    with pytest.raises(Exception):                                                                      # the device encoder, no device: loud
        ref.cv2.imwrite(str(tmp_path / "c.jpg"), img)
    assert not (tmp_path / "c.jpg").exists()
    assert ref.cv2.IMWRITE_JPEG_QUALITY == 1 and ref.cv2.cvtColor is not None                          # everything else is cv2's


def test_finalize_hooks_leave_plain_arrays_to_the_reference():
    from b200mosaic import run as brun
    ref = _ref_module()
    brun.install_device_finalize(ref)
    img = np.zeros((8, 8, 3), np.uint8)
    assert ref.crop_black_areas(img, threshold=80, margin=30) is img and ref.calls == ["crop"]
    assert ref.scale_to_screen(img) is img and ref.calls == ["crop", "scale"]


def test_product_never_imports_the_oracle():
    import pathlib
    pkg = pathlib.Path(__file__).resolve().parent.parent / "real-time-video-mosaic_b200"
    for p in pkg.glob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p.name
