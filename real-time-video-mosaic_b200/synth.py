"""Synthetic drone-sweep generator: a textured ground plane seen through a camera whose
pose drifts by a known homography per frame.  Input generation only (host, cv2/NumPy) -- not part of the hot path."""
from __future__ import annotations

import numpy as np
import cv2


def make_ground(size=4096, seed=1234, n_shapes=None):
    """u8 BGR ground texture: 5 octaves of band-limited noise + random filled shapes, clamped to [1,255]."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((size, size, 3), np.float32)
    for sigma, amp in ((2, 1.0), (4, 0.8), (8, 0.6), (16, 0.4), (32, 0.3)):
        s = max(8, size // sigma)
        n = rng.standard_normal((s, s, 3)).astype(np.float32)
        n = cv2.resize(n, (size, size), interpolation=cv2.INTER_CUBIC)
        acc += amp * n
    acc = (acc - acc.mean()) / (acc.std() + 1e-6)
    img = np.clip(128 + 40 * acc, 1, 255).astype(np.uint8)
    if n_shapes is None:
        n_shapes = int(4000 * (size / 8192.0) ** 2) + 200
    for _ in range(n_shapes):
        kind = rng.integers(0, 3)
        c = tuple(int(v) for v in rng.integers(20, 256, 3))
        x, y = int(rng.integers(0, size)), int(rng.integers(0, size))
        a, b = int(rng.integers(8, 120)), int(rng.integers(8, 120))
        if kind == 0:
            cv2.rectangle(img, (x, y), (x + a, y + b), c, -1)
        elif kind == 1:
            cv2.ellipse(img, (x, y), (a // 2 + 1, b // 2 + 1), float(rng.uniform(0, 180)), 0, 360, c, -1)
        else:
            cv2.line(img, (x, y), (x + a - 60, y + b - 60), c, int(rng.integers(1, 6)))
    return np.maximum(img, 1)


class DroneSweep:
    """Frame t = warp(ground, inv(C_t)) + noise;  C_t = C_{t-1} @ D_t maps frame-t pixels to ground, so the true
    relative homography cur->prev (what findHomography estimates, main.py:727) is exactly D_t."""

    def __init__(self, width=1920, height=1080, seed=1234, ground=None, ground_size=4096, max_step=12.0,
                 noise_sigma=2.0, start=None, max_travel=None):
        self.w, self.h = width, height
        self.rng = np.random.default_rng(seed)
        self.ground = make_ground(ground_size, seed) if ground is None else ground
        gs = self.ground.shape[0]
        self.C = np.eye(3)
        sx, sy = start if start is not None else ((gs - width) * 0.5, gs - height - 64.0)
        self.C[0, 2], self.C[1, 2] = sx, sy
        self.max_step = max_step
        self.noise_sigma = noise_sigma
        self.t = 0
        self.max_travel = max_travel      # reverse the along-track direction after this many px (keeps the mosaic on the canvas)
        self._y0 = sy
        self.dir = np.array([0.0, -1.0])      # fly "up" the ground: the canvas grows upward like the reference's layout
        self._noise = None
        if noise_sigma > 0:
            self._noise = np.clip(self.rng.normal(0, noise_sigma, (height + 64, width + 64, 3)), -127, 127).astype(np.int8)
        self.D_true = []

    def _step_pose(self):
        rng = self.rng
        step = self.max_step * rng.uniform(0.6, 1.0)
        gs = self.ground.shape[0]
        # serpentine: reverse the along-track direction near the ground borders, drift sideways slowly
        cy = self.C[1, 2]
        if self.max_travel is not None and self.dir[1] < 0 and (self._y0 - cy) > self.max_travel:
            self.dir = np.array([0.0, 1.0])
        elif self.max_travel is not None and self.dir[1] > 0 and cy > self._y0 - 8:
            self.dir = np.array([0.0, -1.0])
        elif cy < 96 and self.dir[1] < 0:
            self.dir = np.array([0.0, 1.0])
        elif cy > gs - self.h - 96 and self.dir[1] > 0:
            self.dir = np.array([0.0, -1.0])
        side = rng.normal(0, 0.5)
        tx, ty = self.dir * step + np.array([side, 0.0])
        ang = np.deg2rad(rng.normal(0, 0.03))
        sc = np.exp(rng.normal(0, 2e-4))
        D = np.array([[sc * np.cos(ang), -sc * np.sin(ang), tx],
                      [sc * np.sin(ang), sc * np.cos(ang), ty],
                      [rng.normal(0, 2e-7), rng.normal(0, 2e-7), 1.0]])
        return D

    def next(self):
        if self.t > 0:
            D = self._step_pose()
            self.C = self.C @ D
            self.D_true.append(D)
        frame = cv2.warpPerspective(self.ground, self.C, (self.w, self.h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
        if self._noise is not None:
            ox, oy = int(self.rng.integers(0, 64)), int(self.rng.integers(0, 64))
            n = self._noise[oy:oy + self.h, ox:ox + self.w]
            frame = np.clip(frame.astype(np.int16) + n, 1, 255).astype(np.uint8)
        self.t += 1
        return frame

    def frames(self, n):
        return [self.next() for _ in range(n)]
