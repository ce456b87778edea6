"""Minimal `gymnasium.spaces.Box` stand-in (gymnasium is not part of this image).

Only what the reference touches on the make_env surface (src/pcgym/pcgym.py:68-92, 124):
`.low`, `.high`, `.shape`, `.dtype`, `.sample()`, `.contains()`.  If gymnasium is importable its
Box is used instead so SB3-style wrappers see the real type.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not available in the build image
    from gymnasium.spaces import Box as _GymBox
except Exception:  # noqa: BLE001
    _GymBox = None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        low, high = np.asarray(low), np.asarray(high)
        if shape is not None:
            low, high = np.broadcast_to(low, shape), np.broadcast_to(high, shape)
        self.low = low.astype(dtype, copy=True)
        self.high = high.astype(dtype, copy=True)
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng(seed)

    @property
    def shape(self):
        return self.low.shape

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


def Box(low, high, dtype=np.float64, **kw):
    if _GymBox is not None:  # pragma: no cover
        return _GymBox(low=np.asarray(low, dtype=dtype), high=np.asarray(high, dtype=dtype), dtype=dtype, **kw)
    return _Box(low, high, dtype=dtype, **kw)
