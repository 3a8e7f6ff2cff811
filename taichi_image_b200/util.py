"""Small host-side helpers (reference: util.py:7-84)."""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import List

import numpy as np

cache = lru_cache(maxsize=None)      # util.py:7


@dataclass
class Bounds:                        # util.py:21-47 (host-side mirror; the device reductions live in csrc/)
    min: float
    max: float

    def span(self):
        return self.max - self.min

    def union(self, other):
        self.min = min(self.min, other.min)
        self.max = max(self.max, other.max)

    def expand(self, v):
        self.min = min(self.min, v)
        self.max = max(self.max, v)

    def to_vec(self):
        return np.array([self.min, self.max], dtype=np.float32)

    def scale_range(self, v):
        return (v - self.min) / self.span()


def union_bounds(bounds: List[Bounds]) -> Bounds:      # util.py:63-69
    result = Bounds(np.inf, -np.inf)
    for b in bounds:
        result.union(b)
    return result


def bounds_to_np(b: Bounds):                           # util.py:71-72
    return np.array([b.min, b.max])


def bounds_from_np(b) -> Bounds:                       # util.py:74-75
    return Bounds(float(b[0]), float(b[1]))


def lerp(t, a, b):                                     # util.py:82-84
    return a + t * (b - a)
