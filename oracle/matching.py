"""NumPy restatements of the matchers the reference uses (TEST INFRASTRUCTURE, see oracle/__init__.py).
main.py:686-698: SIFT -> BFMatcher().knnMatch(k=2) + Lowe ratio 0.7; ORB -> BFMatcher(NORM_HAMMING, crossCheck=True).match;
then a stable sort by distance.  Tie rules pinned in SURVEY.md A.6 and against live cv2 in tests/test_oracle_match_cpu.py."""
from __future__ import annotations

import numpy as np

_POP = np.array([bin(i).count("1") for i in range(256)], dtype=np.int32)


def hamming_matrix(q, t):
    return _POP[np.bitwise_xor(q[:, None, :], t[None, :, :])].sum(axis=2)


def match_hamming_crosscheck(des_q, des_t):
    """mutual nearest neighbours, argmin first-index ties in both directions, ordered by queryIdx; then stable sort by
    distance (main.py:698).  Returns (m,3) float64 rows (queryIdx, trainIdx, distance)."""
    d = hamming_matrix(des_q, des_t)
    nn_q = d.argmin(axis=1)
    nn_t = d.argmin(axis=0)
    q = np.arange(len(des_q))
    keep = nn_t[nn_q] == q
    m = np.stack([q[keep], nn_q[keep], d[q[keep], nn_q[keep]]], axis=1).astype(np.float64)
    return m[np.argsort(m[:, 2], kind="stable")]


def match_l2_ratio(des_q, des_t, ratio=0.7):
    """two nearest train rows per query under L2 (ties -> lower train index), keep if d1 < 0.7*d2 evaluated in double on
    the float32 distances (main.py:691), stable sort by distance."""
    a = des_q.astype(np.float64); b = des_t.astype(np.float64)
    d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (a @ b.T)     # exact: integer descriptors < 2^24
    order = np.argsort(d2, axis=1, kind="stable")[:, :2]
    rows = []
    for qi in range(len(des_q)):
        t1, t2 = order[qi]
        d1 = np.sqrt(np.float32(d2[qi, t1])); dd2 = np.sqrt(np.float32(d2[qi, t2]))
        if float(d1) < ratio * float(dd2):
            rows.append((qi, t1, float(d1)))
    m = np.array(rows, dtype=np.float64).reshape(-1, 3)
    return m[np.argsort(m[:, 2], kind="stable")]
