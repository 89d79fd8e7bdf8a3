"""real-time-video-mosaic_b200 (import alias: `b200mosaic`) -- B200-native drop-in for the frame-stitching hot path of
PROcessorI/Real-Time-Video-Mosaic (`VideMosaic.process_frame / findHomography / warp`, main.py:676-977).

Only what the path needs lives here: `csrc/` (hand-written sm_100a CUDA + the C ABI of include/b200mosaic.h),
`_lib.py` (ctypes binding), `mosaic.py` (host-side mirror of the reference class), `ops.py` (stage entry points on
device buffers for parity tests), `synth.py` (synthetic drone-sweep generator for benchmarks)."""
from ._lib import B200MosaicError, load  # noqa: F401
from .mosaic import VideMosaic  # noqa: F401

__all__ = ["VideMosaic", "B200MosaicError", "load"]
