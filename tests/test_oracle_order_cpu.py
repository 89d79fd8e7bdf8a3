"""oracle/cvorder.py (libstdc++ introselect + partition restated) against the REAL std:: algorithms and live cv2's order."""
import numpy as np
import pytest

from oracle import cvorder
from oracle import orb as oorb


@pytest.fixture(scope="module", autouse=True)
def _build_ref():
    from oracle import build_ref
    build_ref.build()


_killer = cvorder.adversarial_input


def test_restatement_equals_real_libstdcxx_random():
    rng = np.random.default_rng(7)
    for n, k, hi in [(1, 1, 5), (3, 2, 5), (4, 1, 3), (10, 3, 4), (50, 50, 9), (51, 50, 9), (500, 304, 60), (500, 152, 10 ** 6),
                     (4096, 700, 200), (20000, 304, 235), (20000, 1, 235), (20000, 19999, 235), (3000, 0, 5), (777, 776, 2)]:
        for rep in range(3):
            r = rng.integers(0, hi, n).astype(np.float32)
            assert np.array_equal(cvorder.retain_best_order(r, k), cvorder.stl_retain_best(r, k)), (n, k, hi)
    r = np.zeros(1000, np.float32)                  # all equal
    assert np.array_equal(cvorder.retain_best_order(r, 10), cvorder.stl_retain_best(r, 10))
    r = np.arange(1000, dtype=np.float32)           # sorted both ways
    assert np.array_equal(cvorder.retain_best_order(r, 100), cvorder.stl_retain_best(r, 100))
    assert np.array_equal(cvorder.retain_best_order(r[::-1], 100), cvorder.stl_retain_best(r[::-1], 100))


def test_heap_select_fallback_equals_real_libstdcxx():
    hit = 0
    for n, nth in [(2000, 1000), (5000, 303), (1500, 1400)]:
        r = _killer(n, nth)
        calls = {"n": 0}
        orig = cvorder._heap_select

        def spy(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)
        cvorder._heap_select = spy
        try:
            got = cvorder.nth_element_order(r, nth)
        finally:
            cvorder._heap_select = orig
        hit += calls["n"]
        assert np.array_equal(got, cvorder.stl_nth_element(r, nth)), (n, nth)
        assert np.array_equal(cvorder.retain_best_order(r, nth + 1), cvorder.stl_retain_best(r, nth + 1))
    assert hit >= 1, "no adversarial input reached the depth limit"


def _orb_in_cv_order(gray, nfeatures=700):
    levels = oorb.build_pyramid(gray)
    scales = oorb.level_scales()
    quotas = oorb.level_quotas(nfeatures)
    out = []
    for l, img in enumerate(levels):
        h, w = img.shape
        xs, ys, sc = oorb.fast_nms(oorb.fast_score_map(img))
        inb = (xs >= oorb.EDGE) & (xs < w - oorb.EDGE) & (ys >= oorb.EDGE) & (ys < h - oorb.EDGE)
        xs, ys, sc = xs[inb], ys[inb], sc[inb]
        o1 = cvorder.stl_retain_best(sc.astype(np.float32), 2 * quotas[l])
        xs, ys = xs[o1], ys[o1]
        if len(xs) == 0:
            continue
        resp = oorb.harris_responses(img, xs, ys)
        o2 = cvorder.stl_retain_best(resp, quotas[l])
        for x, y in zip(xs[o2], ys[o2]):
            out.append((float(np.float32(x) * scales[l]), float(np.float32(y) * scales[l]), float(l)))
    return np.asarray(out).reshape(-1, 3)


def test_orb_order_equals_live_cv2(golden_dir):
    import cv2
    frames = np.load(golden_dir / "clip01_frames.npz")["frames"]
    for t in (0, 3):
        gray = cv2.cvtColor(frames[t], cv2.COLOR_BGR2GRAY)
        kp, _ = oorb.cv_detect_and_compute(gray)
        assert np.array_equal(_orb_in_cv_order(gray), kp[:, [0, 1, 5]])
    rng = np.random.default_rng(5)                 # a corner-rich frame: every level is over quota
    gray = cv2.GaussianBlur(rng.integers(0, 256, (480, 640)).astype(np.uint8), (0, 0), 1.2)
    gray = cv2.normalize(gray, None, 0, 255, cv2.NORM_MINMAX)
    kp, _ = oorb.cv_detect_and_compute(gray)
    assert len(kp) >= 700
    assert np.array_equal(_orb_in_cv_order(gray), kp[:, [0, 1, 5]])


def test_sift_order_equals_live_cv2(golden_dir):
    """SIFT: removeDuplicatedSorted's order, then retainBest(700).  cv2's own angles jitter by an ulp between calls, so the
    comparison is on (pt, octave, response)."""
    import cv2
    rng = np.random.default_rng(11)
    gray = cv2.GaussianBlur(rng.integers(0, 256, (360, 640)).astype(np.uint8), (0, 0), 1.5)
    gray = cv2.normalize(gray, None, 0, 255, cv2.NORM_MINMAX)
    ok = False
    for attempt in range(4):                       # the set itself is not perfectly repeatable either (orientation peaks at 0.8 max)
        kall = cv2.SIFT_create(0).detect(gray, None)
        k700 = cv2.SIFT_create(700).detect(gray, None)
        assert len(kall) > 1400
        xs = np.array([k.pt[0] for k in kall])
        assert np.all(np.diff(xs) >= 0)            # removeDuplicatedSorted leaves them sorted by x first
        resp = np.array([k.response for k in kall], np.float32)
        o = cvorder.stl_retain_best(resp, 700)
        got = [(kall[i].pt, kall[i].octave, kall[i].response) for i in o]
        want = [(k.pt, k.octave, k.response) for k in k700]
        assert np.array_equal(o, cvorder.retain_best_order(resp, 700))
        if got == want:
            ok = True
            break
    assert ok
