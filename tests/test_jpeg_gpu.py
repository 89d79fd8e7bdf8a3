"""mosaic.jpg on the device (csrc/jpeg.cu): the bytes of cv2.imwrite / cv2.imencode('.jpg') -- /root/reference/main.py:1664-1665
writes the screen-scaled mosaic with cv2's defaults -- reproduced byte for byte through the C ABI (bm_jpeg_encode, bm_finalize_jpeg).
The arithmetic is pinned on the CPU in tests/test_oracle_jpeg_cpu.py (oracle/jpeg.py == cv2.imencode)."""
import numpy as np
import cv2
import pytest

from oracle import jpeg as ojpeg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import b200mosaic.ops as o
    return o


@pytest.fixture(scope="module")
def frames(golden_dir):
    return np.load(golden_dir / "clip01_frames.npz")["frames"]


def _img(h, w, kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "extreme":
        yy, xx = np.mgrid[0:h, 0:w]
        return (((yy // 3 + xx // 5) & 1) * 255).astype(np.uint8)[..., None].repeat(3, 2) ^ np.array([0, 255, 0], np.uint8)
    if kind == "black":
        return np.zeros((h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 3 + yy) % 256, (yy * 2 + xx // 2) % 256, (xx + yy * 5) % 256], -1).astype(np.uint8)
    return cv2.GaussianBlur(img, (0, 0), 2.0)


@pytest.mark.parametrize("size", [(1, 1), (8, 8), (16, 16), (17, 23), (100, 75), (121, 200), (240, 427), (30, 1), (1, 40), (480, 854)])
@pytest.mark.parametrize("kind", ["smooth", "noise", "extreme", "black"])
def test_device_jpeg_equals_cv2_and_oracle(ops, size, kind):
    img = _img(size[0], size[1], kind, seed=size[0] * 31 + size[1])
    got = ops.jpeg_encode(img)
    ok, ref = cv2.imencode(".jpg", img)
    assert ok and got == ref.tobytes()
    if size[0] * size[1] <= 121 * 200:
        assert got == ojpeg.encode(img)


@pytest.mark.parametrize("quality", [1, 30, 50, 75, 90, 100])
def test_device_jpeg_other_qualities(ops, quality):
    img = _img(270, 483, "smooth", seed=quality)
    ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok and ops.jpeg_encode(img, quality) == ref.tobytes()


@pytest.mark.parametrize("size", [(1080, 1920), (1079, 1917), (2160, 3840)])
def test_device_jpeg_at_screen_and_4k_sizes(ops, size):
    """main.py:1656 scales the mosaic to the 1920 x 1080 screen; 4K is config 5's frame size.  Photographic content + noise + a black rim."""
    from b200mosaic.synth import make_ground
    g = make_ground(4096, seed=3)[:size[0], :size[1]].copy()
    rng = np.random.default_rng(1)
    g[size[0] // 2:, : size[1] // 3] = rng.integers(0, 256, g[size[0] // 2:, : size[1] // 3].shape, dtype=np.uint8)
    g[:25] = 0
    g[:, -31:] = 0
    ok, ref = cv2.imencode(".jpg", g)
    got = ops.jpeg_encode(g)
    assert ok and len(got) == len(ref) and got == ref.tobytes()


def test_finalize_jpeg_is_the_file_the_reference_writes(frames):
    """main.py:1647-1666: crop_black_areas -> scale_to_screen -> cv2.imwrite('mosaic.jpg').  One device pass, only the file comes back."""
    import b200mosaic
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    for t in range(1, len(frames)):
        vm.process_frame(frames[t], t)
    data = vm.finalize_jpeg(threshold=80, margin=30)
    scaled = vm.finalize(threshold=80, margin=30)
    ok, ref = cv2.imencode(".jpg", scaled)
    assert ok and data == ref.tobytes()
    assert vm.last_final_size == (scaled.shape[1], scaled.shape[0])
    assert np.array_equal(cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR), cv2.imdecode(ref, cv2.IMREAD_COLOR))
    small = vm.finalize_jpeg(threshold=80, margin=30, target_w=640, target_h=360, quality=80)
    ok, ref = cv2.imencode(".jpg", vm.finalize(threshold=80, margin=30, target_w=640, target_h=360), [cv2.IMWRITE_JPEG_QUALITY, 80])
    assert ok and small == ref.tobytes()


def test_jpeg_buffer_too_small_is_reported(ops):
    import ctypes as C
    from b200mosaic import _lib
    lib = _lib.load()
    img = np.ascontiguousarray(_img(64, 64, "noise"))
    out = np.empty(700, np.uint8); n = C.c_size_t(0)
    st = lib.bm_jpeg_encode(img.ctypes.data_as(C.c_void_p), 64, 64, 95, 0, out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(n))
    assert st != 0 and n.value == len(cv2.imencode(".jpg", img)[1])
