"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/b200mosaic.h declares;
no compute is attempted without a GPU, and the product fails loudly (no CPU fallback) when CUDA is missing."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "b200mosaic.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import b200mosaic
    lib = b200mosaic.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200mosaic.h but not exported"
    assert lib.bm_version() >= 100


def test_python_mirror_has_reference_surface():
    import b200mosaic
    import inspect
    sig = inspect.signature(b200mosaic.VideMosaic.__init__)
    names = list(sig.parameters)[1:8]
    assert names == ["first_image", "output_height_times", "output_width_times", "detector_type",
                     "show_intermediate", "output_dir", "visualize"]          # main.py:17
    assert sig.parameters["output_height_times"].default == 2
    assert sig.parameters["output_width_times"].default == 1.2
    assert sig.parameters["detector_type"].default == "sift"
    for m in ("process_frame", "match", "findHomography", "warp", "validate_homography", "smooth_homography",
              "get_transformed_corners", "draw_border"):
        assert hasattr(b200mosaic.VideMosaic, m), m


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    import b200mosaic
    frame = np.full((64, 96, 3), 7, np.uint8)
    with pytest.raises(b200mosaic.B200MosaicError):
        b200mosaic.VideMosaic(frame, detector_type="orb", visualize=False)


def test_product_does_not_import_oracle():
    pkg = ROOT / "real-time-video-mosaic_b200"
    for p in list(pkg.glob("*.py")) + list(pkg.glob("csrc/*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".h"):
            t = p.read_text(errors="ignore")
            assert "import oracle" not in t and "from oracle" not in t, p


def test_validate_homography_single_native_implementation():
    """bm_validate_homography (host-only) == the reference's method on accept / every reject branch / the NaN quirk
    (main.py:761-801); the Python mirror and sharding.py route through it."""
    import contextlib
    import io
    import b200mosaic
    from b200mosaic import _lib, sharding
    from oracle.mosaic_ref import RefMosaic
    frame = np.zeros((64, 96, 3), np.uint8); frame[8:40, 8:60] = 200
    ref = RefMosaic(frame, detector_type="orb")
    rng = np.random.default_rng(1)
    cases = [np.eye(3)]
    for _ in range(200):
        H = np.eye(3)
        H[:2, :2] += rng.normal(size=(2, 2)) * rng.choice([0.01, 0.2, 0.6])
        H[:2, 2] = rng.normal(size=2) * rng.choice([5.0, 40.0, 80.0])
        H[2, :2] = rng.normal(size=2) * rng.choice([1e-5, 8e-4, 3e-3])
        cases.append(H)
    neg = np.eye(3); neg[0, 0] = -1.0                      # det(H[:2,:2]) < 0 -> sqrt = NaN -> comparisons False -> PASSES
    nan = np.eye(3); nan[1, 1] = np.nan
    inf = np.eye(3); inf[0, 2] = np.inf
    cases += [neg, nan, inf]
    reasons = set()
    for H in cases:
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(invalid="ignore"):
            want = bool(ref.validate_homography(H))
        r, _v = _lib.validate_homography(H)
        reasons.add(r)
        assert (r == _lib.BM_VAL_OK) == want, (H, r)
        assert sharding.validate_homography(H) == want
    assert reasons == {_lib.BM_VAL_OK, _lib.BM_VAL_NAN, _lib.BM_VAL_TRANSLATION, _lib.BM_VAL_SCALE, _lib.BM_VAL_PERSPECTIVE}
    assert _lib.validate_homography(neg)[0] == _lib.BM_VAL_OK
    assert _lib.validate_homography(None)[0] == _lib.BM_VAL_NAN


def test_library_carries_its_source_digest():
    """lib/stamp.txt (next to the library: lib/ travels to the GPU box, csrc/_obj/ does not) vouches for the binary: after load() the
    stamp equals the digest of csrc + the public header + the nvcc flags, so a box that received the tree loads the library as is and a
    stale binary is rebuilt (once: builders hold lib/.build.lock)"""
    import importlib
    import b200mosaic
    b200mosaic.load()
    bld = importlib.import_module("real-time-video-mosaic_b200.build")
    srcs = sorted(bld.CSRC.glob("*.cu"))
    hdrs = sorted(bld.CSRC.glob("*.cuh")) + sorted(bld.CSRC.glob("*.h")) + [bld.HERE.parent / "include" / "b200mosaic.h"]
    assert bld.STAMP.parent == bld.LIB.parent
    assert bld._up_to_date(bld._digest(srcs + hdrs))
    assert not bld._up_to_date("something else")
