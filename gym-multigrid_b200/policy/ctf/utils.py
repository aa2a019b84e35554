"""Grid helpers of the CtF heuristic policies (reference: policy/ctf/utils.py:17-136, utils/map.py:42-61).

`a_star` returns, cell for cell, the path the reference's implementation returns - including its tie-breaking, which is
what decides the red agents' moves and so has to be reproduced, not merely "a shortest path":

  * the frontier is ordered by the whole node record `(f, g, h, parent, loc)`, compared as nested tuples
    (utils.py:9-14 is a NamedTuple on a heapq) - equal `f` falls through to `g`, `h`, then the parent chain, then `loc`;
  * neighbours are visited in the order (0,+1), (0,-1), (+1,0), (-1,0) (utils.py:64) and a cell already on the frontier
    (or already expanded) is replaced only by a STRICTLY smaller `f` (utils.py:96-116) - the first discoverer wins ties;
  * a cell blocks only if the map holds the value 8 there (utils.py:73).  CtF maps use 6 for obstacles
    (core/world.py:66-79), so on them nothing blocks and paths run through obstacles; kept as is.

The reference finds "is this cell on the frontier / expanded" by scanning Python lists (quadratic); here both are
dictionaries keyed by cell and replaced frontier records are dropped lazily when they surface, which pops the same
records in the same order (the heap minimum under a total order does not depend on how the heap is stored)."""
from __future__ import annotations

from heapq import heappop, heappush

import numpy as np

NEIGHBOUR_ORDER = ((0, 1), (0, -1), (1, 0), (-1, 0))   # utils.py:64
BLOCKING_VALUE = 8                                     # utils.py:73


def manhattan_distance(p1, p2) -> int:
    """utils.py:123-136"""
    return abs(p1[0] - p2[0]) + abs(p1[1] - p2[1])


def _cell(p):
    return int(p[0]), int(p[1])


def a_star(start, end, map) -> list:
    """Cells from `start` to `end` inclusive ([] when `end` cannot be reached); `map` is indexed `[p[0]][p[1]]`."""
    start, end = _cell(start), _cell(end)
    blocked = (np.asarray(map) == BLOCKING_VALUE)
    rows, cols = blocked.shape
    blocked = blocked.tolist()
    h0 = manhattan_distance(start, end)
    first = (h0, 0, h0, None, start)
    frontier = [first]                   # heap of node records; may hold records that were replaced since
    live = {start: first}                # cell -> its current frontier record
    expanded = {}                        # cell -> the record it was expanded with
    while frontier:
        node = heappop(frontier)
        cell = node[4]
        if live.get(cell) is not node:
            continue                     # a replaced record surfacing late
        del live[cell]
        expanded[cell] = node
        if cell == end:
            path = []
            while node is not None:
                path.append(node[4])
                node = node[3]
            path.reverse()
            return path
        g = node[1] + 1
        for dx, dy in NEIGHBOUR_ORDER:
            nxt = (cell[0] + dx, cell[1] + dy)
            if not (0 <= nxt[0] < rows and 0 <= nxt[1] < cols) or blocked[nxt[0]][nxt[1]]:
                continue
            h = manhattan_distance(nxt, end)
            f = g + h
            seen = expanded.get(nxt)
            if seen is not None:
                if f >= seen[0]:
                    continue
                del expanded[nxt]        # utils.py:96-103: back onto the frontier
            else:
                seen = live.get(nxt)
                if seen is not None and f >= seen[0]:
                    continue
            rec = (f, g, h, node, nxt)
            live[nxt] = rec
            heappush(frontier, rec)
    return []


def position_in_positions(position, positions) -> bool:
    """utils/map.py:42-53"""
    return any(position[0] == p[0] and position[1] == p[1] for p in positions)


def closest_area_pos(pos, area):
    """utils/map.py:56-61: the FIRST element of `area` at minimal Euclidean distance from `pos`.  The reference takes the
    argmin of float norms; squared integer distances order (and tie) identically, so the choice is the same.
    An empty `area` raises ValueError as `np.argmin([])` does there."""
    cells = np.asarray(area, dtype=np.int64).reshape(-1, 2)
    if cells.shape[0] == 0:
        raise ValueError("attempt to get argmin of an empty sequence")
    d = cells - np.asarray(_cell(pos), dtype=np.int64)
    return area[int(np.argmin((d * d).sum(axis=1)))]
