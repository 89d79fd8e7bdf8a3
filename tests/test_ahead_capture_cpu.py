"""run.AheadCapture (the launcher's reader thread around cv2.VideoCapture) on the CPU: without a CUDA device the pinned
ring falls back to pageable buffers, the read / peek protocol is the same -- the unmodified driver loop of main.py:1597-1613 must see exactly
the frames cv2.VideoCapture decodes, in order, and the end of the stream."""
import zlib

import numpy as np


def _direct(path, limit):
    import cv2
    cap = cv2.VideoCapture(path)
    out = []
    while len(out) < limit:
        ok, f = cap.read()
        if not ok:
            break
        out.append(zlib.crc32(f.tobytes()))
    cap.release()
    return out


def test_read_and_peek_follow_videocapture(golden_dir):
    import cv2
    from b200mosaic import run as brun
    path = str(golden_dir / "clip01.mp4")
    want = _direct(path, 40)
    brun.AheadCapture.real = cv2.VideoCapture
    cap = brun.AheadCapture(path)
    assert cap.isOpened()                                   # attribute pass-through to the real capture
    got = []
    for t in range(40):
        ok, f = cap.read()
        assert ok and f.dtype == np.uint8 and f.shape == (480, 854, 3)
        crc = zlib.crc32(f.tobytes())
        if t % 3 == 0:                                      # what the swapped process_frame does: look at the next three frames
            nxt = cap.peek_next(3)
            assert len(nxt) == 3 and all(x is not None for x in nxt)
            peeked = [zlib.crc32(x.tobytes()) for x in nxt]
            assert zlib.crc32(f.tobytes()) == crc           # peeking never touches the frame handed out (ring of 12 buffers)
            if t + 3 < 40:
                assert peeked == want[t + 1:t + 4]
        got.append(crc)
    assert got == want
    cap.release()


def test_end_of_stream(golden_dir):
    import cv2
    from b200mosaic import run as brun
    path = str(golden_dir / "clip01.mp4")
    brun.AheadCapture.real = cv2.VideoCapture
    cap = brun.AheadCapture(path)
    n = 0
    last_peek = None
    while True:
        ok, f = cap.read()
        if not ok:
            break
        n += 1
        if n >= 588:
            last_peek = cap.peek_next(3)                    # past the end: None in place of frames that do not exist
    assert n == 592 and last_peek == [None, None, None]
    cap.release()
