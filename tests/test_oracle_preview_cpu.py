"""The preview oracle (oracle/preview.py) against the committed fixture of the reference's call sequence and against live Pillow."""
import hashlib

import numpy as np
import pytest
from pathlib import Path

from oracle import preview as opv

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "preview.npz"), np.load(GOLD / "clip01_orb.npz")["canvas_final"]


def test_restatement_matches_reference_calls_fixture(gold):
    g, canvas = gold
    assert np.array_equal(opv.thumbnail(canvas), g["thumb"])
    assert np.array_equal(opv.thumbnail(g["small_in"]), g["small_thumb"])
    assert np.array_equal(opv.thumbnail(canvas, size=(1100, 200)), g["mixed_thumb"])
    up = opv.thumbnail(canvas, size=(640, 600))
    assert np.array_equal(np.frombuffer(hashlib.sha256(up.tobytes()).digest(), np.uint8), g["up_sha256"])
    assert np.array_equal(up[::4, ::4], g["up_sub4"])


def test_restatement_matches_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for (h, w, ow, oh) in [(480, 640, 400, 300), (100, 130, 400, 300), (301, 403, 400, 300), (300, 400, 400, 300), (37, 900, 400, 300),
                           (720, 768, 33, 17)]:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        im[: h // 3] = 0
        im[h // 2:, w // 2:] = 255                         # saturated plateau: the negative lobes must clamp, not wrap
        want = np.asarray(Image.fromarray(im).resize((ow, oh)))
        assert np.array_equal(opv.pil_resize_bicubic(im, (ow, oh)), want), (h, w, ow, oh)


def test_gui_call_sequence_equals_restatement(gold):
    pytest.importorskip("PIL.Image")
    g, canvas = gold
    assert np.array_equal(opv.gui_thumbnail(canvas.astype(np.float64)), opv.thumbnail(canvas))     # the reference canvas is float64


def test_table_shape_and_normalisation():
    b, kk = opv.axis_table(2304, 400)
    assert kk.shape == (400, 25) and b[:, 0].min() == 0 and (b[:, 0] + b[:, 1]).max() == 2304
    assert np.abs(kk.sum(axis=1) - (1 << opv.PRECISION_BITS)).max() <= 13          # rounding of <= 25 coefficients
