"""Observation / action space objects.  Uses gymnasium's when it is installed, else duck-typed stand-ins with the
attributes the reference and SB3 read (``.n``, ``.shape``, ``.dtype``, ``.low``, ``.high``, ``sample``, ``contains``).
Reference: ``envs/CubicEnv.py:56-62``, ``envs/simpleEnv.py:55-67``."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium import spaces as _gspaces
    Discrete = _gspaces.Discrete
    Box = _gspaces.Box
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Discrete:  # type: ignore[no-redef]
        def __init__(self, n: int, seed=None):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return int(self._rng.integers(0, self.n))

        def contains(self, x) -> bool:
            try:
                return 0 <= int(x) < self.n
            except (TypeError, ValueError):
                return False

        def __repr__(self):
            return f"Discrete({self.n})"

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def cubic_spaces():
    """``Discrete(6)`` and ``Box(-1, 1, (80,), float32)`` (``CubicEnv.py:56-62``)."""
    return Discrete(6), Box(low=np.full(80, -1.0, dtype=np.float32), high=np.full(80, 1.0, dtype=np.float32),
                            dtype=np.float32)


def simple_spaces(local_map_length: int):
    """``simpleEnv.py:55-67``: 6*L ray cells in [-1, 2], 6 distances in [0, inf), last action in [0, 5]."""
    L = int(local_map_length)
    low = np.array([-1] * (6 * L) + [0] * 6 + [0], dtype=np.float32)
    high = np.array([2] * (6 * L) + [float("inf")] * 6 + [5], dtype=np.float32)
    return Discrete(6), Box(low=low, high=high, dtype=np.float32)
