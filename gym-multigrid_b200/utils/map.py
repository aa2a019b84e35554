"""`gym_multigrid.utils.map` under this package's name (utils/map.py:7-61): the text-map loader the CtF / Maze constructors and
the reference's tests use (tests/test_ctf.py:16), and the distance helpers behind the info dicts and the scripted policies.
The batched envs compute the same distances on the device (`mg_map_info`); these are for host-side callers."""
from __future__ import annotations

import numpy as np

from ..policy.ctf.utils import closest_area_pos, position_in_positions  # noqa: F401  (utils/map.py:42-61)


def load_text_map(map_path) -> np.ndarray:
    """utils/map.py:22-39: `np.loadtxt(map_path).T`, i.e. field_map[x, y]; arrays are passed through."""
    if isinstance(map_path, (str, bytes)) or hasattr(map_path, "__fspath__"):
        return np.loadtxt(map_path).T
    return np.asarray(map_path)


def distance_points(p1, p2, is_defeated: bool = False) -> float:
    """utils/map.py:7-13: Euclidean distance of two cells, inf for a defeated agent."""
    if is_defeated:
        return float("inf")
    return float(np.sqrt(float((int(p1[0]) - int(p2[0])) ** 2 + (int(p1[1]) - int(p2[1])) ** 2)))


def distance_area_point(point, area) -> float:
    """utils/map.py:16-19: distance from `point` to the closest cell of `area` (ValueError on an empty area, as np.min there)."""
    cells = np.asarray(area, dtype=np.int64).reshape(-1, 2)
    if cells.shape[0] == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    d = cells - np.array([int(point[0]), int(point[1])], dtype=np.int64)
    return float(np.sqrt(float((d * d).sum(axis=1).min())))
