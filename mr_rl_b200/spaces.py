"""Minimal ``gym.spaces.Box`` stand-in (the reference imports gym only for this class,
MR_env.py:7,34-45).  float32 bounds, uniform ``sample`` and a bounds-only ``contains``."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32, seed=None):
        self.low = np.asarray(low).astype(dtype)
        self.high = np.asarray(high).astype(dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.RandomState(seed)

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
