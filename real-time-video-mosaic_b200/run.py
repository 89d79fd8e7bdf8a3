"""Launcher: run the UNMODIFIED reference driver (`main.main`, main.py:1512-1717) with `VideMosaic` swapped for the
B200 implementation.

    python -m b200mosaic.run <video> [--reference-dir DIR] [--output-dir D] [--detector sift|orb]

`--reference-dir` is the checkout that holds the reference's main.py (it is imported, never modified).  YOLO detection,
A* navigation and the mosaic.jpg writing stay on the reference's own code path; `crop_black_areas` + `scale_to_screen` of the
final canvas (main.py:1647-1659) run on the device (`bm_finalize`) so that only the screen-sized image is copied back."""
from __future__ import annotations

import argparse
import importlib.util
import sys
import types
from pathlib import Path


def _stub_missing(name, attrs=()):
    try:
        __import__(name)
    except Exception:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            n = ".".join(parts[:i])
            sys.modules.setdefault(n, types.ModuleType(n))
        for a in attrs:
            setattr(sys.modules[name], a, object)


def load_reference_main(ref_dir: Path):
    # out-of-scope imports of main.py that may be absent on a GPU box; both are only used by detection / navigation
    try:
        import ultralytics  # noqa: F401
    except Exception:
        u = types.ModuleType("ultralytics")

        class _YOLO:
            def __init__(self, *a, **k):
                raise RuntimeError("ultralytics unavailable")      # caught at main.py:45-47, 68-70
        u.YOLO = _YOLO
        sys.modules["ultralytics"] = u
    _stub_missing("pathfinding.core.grid", ("Grid",))
    _stub_missing("pathfinding.core.diagonal_movement", ("DiagonalMovement",))
    _stub_missing("pathfinding.finder.a_star", ("AStarFinder",))
    spec = importlib.util.spec_from_file_location("main", str(ref_dir / "main.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["main"] = mod
    spec.loader.exec_module(mod)
    return mod


class LazyCanvas:
    """What `video_mosaic.output_img` evaluates to under this launcher: a stand-in for the (Hc, Wc, 3) uint8 canvas that only
    copies it from the device when somebody really needs the pixels (`np.asarray`, `.copy()`, `.astype()`, indexing,
    arithmetic).  main()'s finalisation `crop_black_areas(video_mosaic.output_img, ...)` -> `scale_to_screen(cropped)`
    (main.py:1647-1659) recognises it and is served by bm_finalize on the device, so the full canvas (3.2 GB in config 5) never
    crosses PCIe for mosaic.jpg."""
    materialized = 0            # how many times the full canvas was copied to the host through a LazyCanvas (tests read it)

    def __init__(self, vm, getter):
        self._vm, self._getter = vm, getter

    def _arr(self):
        LazyCanvas.materialized += 1
        return self._getter(self._vm)

    shape = property(lambda self: tuple(self._vm._shape))
    dtype = property(lambda self: __import__("numpy").dtype("uint8"))
    ndim = 3
    size = property(lambda self: self.shape[0] * self.shape[1] * self.shape[2])

    def __array__(self, dtype=None, copy=None):
        a = self._arr()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]

    def copy(self):
        return self._arr().copy()

    def astype(self, *a, **k):
        return self._arr().astype(*a, **k)

    def __getitem__(self, i):
        return self._arr()[i]

    def __truediv__(self, o):
        return self._arr() / o

    def __mul__(self, o):
        return self._arr() * o


class AheadCapture:
    """cv2.VideoCapture with a one-frame-ahead reader thread (the reference's loop is
    `ret, frame = cap.read(); video_mosaic.process_frame(frame, n)`, main.py:1597-1613).  Frames are decoded by the real
    cv2.VideoCapture on a background thread into a ring of pinned host buffers; `read()` hands out frame t while frames t+1 .. t+3 are
    already decoded, and `peek_next(3)` lets the swapped `process_frame` stage them (H2D + detect-ahead) while t is processed --
    the unmodified driver loop gets the double-buffered ingest without passing `next_frame`."""
    current = None
    RING = 12                   # pinned buffers: 3 queued + 1 being decoded + 3 peeked + the frames the pipeline still holds

    def __init__(self, *args, **kw):
        import queue
        import threading
        self._cap = AheadCapture.real(*args, **kw)
        self._q = queue.Queue(maxsize=3)
        self._peeked = []
        self._ring, self._slot, self._lib = [], 0, None
        self._stop = False
        self._thread = threading.Thread(target=self._worker, daemon=True)
        self._started = False
        AheadCapture.current = self

    def _buffer(self, shape):
        import ctypes as C
        import numpy as np
        if not self._ring:
            from . import _lib
            self._lib = _lib.load()
            n = int(np.prod(shape))
            for _ in range(self.RING):
                p = C.c_void_p()
                if self._lib.bm_alloc_pinned(n, C.byref(p)) != 0:
                    self._ring.append(np.empty(shape, np.uint8))           # pageable fallback: the library stages it itself
                else:
                    self._ring.append(np.ctypeslib.as_array((C.c_uint8 * n).from_address(p.value)).reshape(shape))
        b = self._ring[self._slot]
        self._slot = (self._slot + 1) % self.RING
        return b

    def _worker(self):
        while not self._stop:
            ok, f = self._cap.read()
            if ok:
                buf = self._buffer(f.shape)
                buf[...] = f
                f = buf
            self._q.put((ok, f))
            if not ok:
                break

    def _pull(self):
        if not self._started:
            self._started = True
            self._thread.start()
        return self._q.get()

    def read(self):
        return self._peeked.pop(0) if self._peeked else self._pull()

    def peek_next(self, k=1):
        """the frames the next k (<= 3) read() calls will return (blocks until they are decoded); None in place of frames past the
        end of the stream"""
        while len(self._peeked) < k and not (self._peeked and not self._peeked[-1][0]):
            self._peeked.append(self._pull())
        out = [f if ok else None for ok, f in self._peeked[:k]]
        return out + [None] * (k - len(out))

    def release(self):
        self._stop = True
        try:
            while True:
                self._q.get_nowait()
        except Exception:
            pass
        self._cap.release()

    def __getattr__(self, name):            # isOpened, get, set, ...
        return getattr(self._cap, name)


def make_swapped_class(det, ahead=True):
    import b200mosaic

    class _Swapped(b200mosaic.VideMosaic):
        def __init__(self, first_image, *args, **kw):
            if det is not None:
                kw["detector_type"] = det   # main() hard-codes "sift" (main.py:1603)
            kw.setdefault("visualize", False)
            super().__init__(first_image, *args, **kw)

        def process_frame(self, frame_cur, frame_count=0, next_frame=None, next2_frame=None, next3_frame=None):
            if next_frame is None and ahead and AheadCapture.current is not None:
                next_frame, next2_frame, next3_frame = AheadCapture.current.peek_next(3)
            return super().process_frame(frame_cur, frame_count, next_frame=next_frame, next2_frame=next2_frame, next3_frame=next3_frame)

        @property
        def output_img(self):
            return LazyCanvas(self, b200mosaic.VideMosaic.output_img.fget)

        @output_img.setter
        def output_img(self, img):
            b200mosaic.VideMosaic.output_img.fset(self, img)

    return _Swapped


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("video")
    ap.add_argument("--reference-dir", default=".")
    ap.add_argument("--output-dir", default=None)
    ap.add_argument("--detector", default="sift", choices=["sift", "orb"])
    ap.add_argument("--no-read-ahead", action="store_true", help="keep cv2.VideoCapture as is (no reader thread)")
    a = ap.parse_args(argv)
    import cv2
    import b200mosaic
    ref = load_reference_main(Path(a.reference_dir))
    try:                                    # headless OpenCV wheels raise in highgui calls (main.py:1616 calls waitKey)
        cv2.waitKey(1)
    except cv2.error:
        cv2.waitKey = lambda *x, **k: -1
        cv2.destroyAllWindows = cv2.imshow = cv2.namedWindow = lambda *x, **k: None
    b200mosaic.VideMosaic.reference_class = ref.VideMosaic
    if not a.no_read_ahead:
        AheadCapture.real = cv2.VideoCapture
        cv2.VideoCapture = AheadCapture    # main() looks cv2.VideoCapture up at call time (main.py:1579)
    swapped = make_swapped_class(a.detector, ahead=not a.no_read_ahead)
    ref.VideMosaic = swapped
    install_device_finalize(ref)
    ref.main(video_path=a.video, show_intermediate=False, output_dir=a.output_dir)


def install_device_finalize(ref):
    """main() calls  cropped = crop_black_areas(video_mosaic.output_img, threshold=80, margin=30)  and then
    scaled = scale_to_screen(cropped)  (main.py:1647-1659).  Under this launcher `video_mosaic.output_img` is a LazyCanvas:
    crop_black_areas recognises it and runs bm_finalize on the device WITHOUT fetching the canvas; it returns a zero-copy
    placeholder of the cropped SHAPE (main() only prints it) that carries the device result, scale_to_screen returns that result.
    Any other use of the two functions (plain ndarrays, explicit target sizes) goes to the reference's implementation; a FAILURE of the
    device path raises -- there is no silent host fallback."""
    import numpy as np
    ref_crop, ref_scale = ref.crop_black_areas, ref.scale_to_screen

    class _Placeholder(np.ndarray):
        pass

    def crop_black_areas(image, threshold=15, margin=5):
        if isinstance(image, LazyCanvas):
            out = image._vm.finalize(threshold, margin)
            x, y, w, h = image._vm.last_crop_rect
            ph = np.lib.stride_tricks.as_strided(np.zeros(1, np.uint8), shape=(h, w, 3), strides=(0, 0, 0)).view(_Placeholder)
            ph._b200_result = out
            return ph
        return ref_crop(image, threshold, margin)

    def scale_to_screen(image, target_w=None, target_h=None):
        if isinstance(image, _Placeholder) and target_w is None and target_h is None and getattr(image, "_b200_result", None) is not None:
            return image._b200_result
        if isinstance(image, LazyCanvas):
            image = np.asarray(image)
        return ref_scale(image, target_w, target_h)

    ref.crop_black_areas, ref.scale_to_screen = crop_black_areas, scale_to_screen
    install_device_imwrite(ref)


def install_device_imwrite(ref):
    """main() ends with cv2.imwrite(os.path.join(output_dir, 'mosaic.jpg'), scaled_mosaic) (main.py:1664-1665; navigation_map.jpg at
    :1696-1697 likewise).  Inside the reference module only, `cv2` becomes a pass-through proxy whose imwrite encodes 3-channel 8-bit
    images going to .jpg / .jpeg with default parameters on the device (`ops.jpeg_encode`: the same bytes libjpeg writes) and hands
    everything else -- other formats, explicit parameters -- to the real cv2.imwrite.  A failure of the device encoder raises (no
    silent host fallback); an unwritable path returns False like cv2.imwrite does."""
    import numpy as np
    from . import ops
    real = ref.cv2

    class _Cv2Proxy:
        def __getattr__(self, name):
            return getattr(real, name)

        @staticmethod
        def imwrite(path, img, params=None):
            if (params is None and isinstance(img, np.ndarray) and img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
                    and img.size > 0 and str(path).lower().endswith((".jpg", ".jpeg"))):
                data = ops.jpeg_encode(img)
                try:
                    with open(path, "wb") as f:
                        f.write(data)
                    return True
                except OSError:
                    return False                         # cv2.imwrite reports an unwritable path the same way
            return real.imwrite(path, img) if params is None else real.imwrite(path, img, params)

    ref.cv2 = _Cv2Proxy()


if __name__ == "__main__":
    main()
