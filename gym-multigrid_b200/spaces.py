"""Action / observation space descriptions.

`gymnasium` is used when it is importable; otherwise these light stand-ins carry the same
attributes the reference exposes (multigrid.py:66,91-112): `Discrete.n`, `Box.shape/low/high/dtype`,
`MultiDiscrete.nvec`.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is not part of the build image
    from gymnasium.spaces import Box, Discrete, MultiDiscrete  # type: ignore  # noqa: F401
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Discrete:
        def __init__(self, n: int, start: int = 0):
            self.n, self.start, self.shape, self.dtype = int(n), int(start), (), np.dtype(np.int64)

        def contains(self, x) -> bool:
            return self.start <= int(x) < self.start + self.n

        def sample(self, rng=None):
            rng = rng or np.random.default_rng()
            return int(rng.integers(self.start, self.start + self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete:
        def __init__(self, nvec, dtype=np.int64):
            self.nvec = np.asarray(nvec, dtype=dtype)
            self.shape, self.dtype = self.nvec.shape, np.dtype(dtype)

        def sample(self, rng=None):
            rng = rng or np.random.default_rng()
            return (rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape, self.dtype = tuple(int(s) for s in shape), np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
