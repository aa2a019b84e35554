"""Tables that let the CtF kernel side decide for the scripted opponents (`mg_set_red_policies`, csrc/policy_kernels.cu).

The reference's `DestinationPolicy.act` (heuristic.py:125-177) needs two things per decision: a target and the first cell of
the A* route towards it.  The route depends on (cell, target) and the map alone (utils.py:17-120), so this module runs the
reference-pinned host A* (`utils.a_star`) once per pair and tabulates the resulting action; targets that are a function of
the cell alone (PatrolPolicy off the border) are tabulated too.  What remains per step - closest blue agent, "is a blue agent
on red ground", the random draws - is done by the kernel.  Host code, numpy only."""
from __future__ import annotations

import numpy as np

from .heuristic import CapturePolicy, FightPolicy, PatrolFightPolicy, PatrolPolicy, RwPolicy
from .utils import BLOCKING_VALUE, a_star, closest_area_pos

KIND = {RwPolicy: 0, FightPolicy: 1, CapturePolicy: 2, PatrolPolicy: 3, PatrolFightPolicy: 4}     # MG_POLICY_* (multigrid_b200.h)
_ACTION_OF_STEP = {(0, 0): 0, (0, -1): 1, (-1, 0): 2, (0, 1): 3, (1, 0): 4}                          # CtfActions, heuristic.py:160-170
MAX_CELLS = 1024


def first_move_table_py(field_map) -> np.ndarray:
    """u8 [cells, cells]: action of the first move of `a_star(start, target, field_map)`, start-major, cell = x * cols + y
    (`stay` when start == target).  A pair without a route gets the action towards the target itself when it is adjacent and
    `stay` otherwise - the reference raises "Invalid direction" there (heuristic.py:144-172); CtF maps have no such pair,
    only the value 8 blocks (utils.py:73).  Pure-Python statement of the table (one `utils.a_star` run per pair); the library's
    `mg_astar_first_moves` computes the same table natively and is what `first_move_table` calls."""
    fm = np.asarray(field_map)
    rows, cols = fm.shape
    cells = rows * cols
    out = np.zeros((cells, cells), np.uint8)
    for s in range(cells):
        start = divmod(s, cols)
        for t in range(cells):
            path = a_star(start, divmod(t, cols), fm)
            nxt = path[1] if len(path) > 1 else divmod(t, cols)
            out[s, t] = _ACTION_OF_STEP.get((nxt[0] - start[0], nxt[1] - start[1]), 0)
    return out


def first_move_table(field_map) -> np.ndarray:
    """The same table from the library's host-side A* (`mg_astar_first_moves`, csrc/astar_host.cu: the same frontier order
    in C++, start cells spread over the host threads): 10x10 in 0.04 s instead of 0.7 s, 32x32 (10^6 routes) in seconds."""
    import ctypes as C

    from ... import _lib
    fm = np.asarray(field_map)
    rows, cols = fm.shape
    blocked = np.ascontiguousarray(fm == BLOCKING_VALUE, dtype=np.uint8)
    out = np.zeros((rows * cols, rows * cols), np.uint8)
    if _lib.load().mg_astar_first_moves(C.c_void_p(blocked.ctypes.data), rows, cols, C.c_void_p(out.ctypes.data)) != 0:
        raise ValueError("mg_astar_first_moves: map too large")
    return out


def build_tables(policies, field_map) -> dict:
    """`policies`: one entry per red agent - None / RwPolicy / FightPolicy / CapturePolicy / PatrolPolicy / PatrolFightPolicy
    objects of this package (exact types: a subclass may decide differently), ego_agent "red", on the env's map."""
    fm = np.asarray(field_map)
    if fm.ndim != 2 or fm.shape[0] != fm.shape[1]:
        raise ValueError("square field maps only")
    S = fm.shape[0]
    cells = S * S
    if cells > MAX_CELLS:
        raise ValueError(f"device policies tabulate A* per (cell, target) pair: maps up to {MAX_CELLS} cells")
    kind, randomness, patrol = [], [], None
    for p in policies:
        if p is None:
            p = RwPolicy()
        if type(p) not in KIND:
            raise TypeError(f"{type(p).__name__} has no device form; use set_enemy_policies(..., device=False)")
        kind.append(KIND[type(p)])
        randomness.append(float(getattr(p, "randomness", 0.0)))
        if kind[-1] == 0:
            continue
        if getattr(p, "ego_agent", "red") != "red":
            raise ValueError("the scripted opponents of CtFMvNEnv play red")
        if p.field_map is not None and not np.array_equal(np.asarray(p.field_map), fm):
            raise ValueError("policy.field_map differs from the env's map")
        if kind[-1] >= 3:
            if not p.border:
                raise ValueError("patrol policy without a border (built without field_map?): the reference fails on its first decision")
            sig = ([tuple(int(v) for v in c) for c in p.border], p._along_border.tolist())
            if patrol is not None and patrol != sig:
                raise ValueError("patrol policies with different borders")
            patrol = sig
    t = dict(kind=np.array(kind, np.int32), randomness=np.array(randomness, np.float64), first_move=first_move_table(fm),
             patrol_goal=np.zeros(cells, np.uint16), on_border=np.zeros(cells, np.uint8), along_border=np.zeros(0, np.uint16))
    if patrol is not None:
        border, along = patrol
        if not along:
            raise ValueError("no border cell has a border neighbour: the reference raises on the border (heuristic.py:333)")
        for c in range(cells):
            x, y = closest_area_pos(divmod(c, S), border)
            t["patrol_goal"][c] = x * S + y
        for x, y in border:
            t["on_border"][x * S + y] = 1
        t["along_border"] = np.array([x * S + y for x, y in along], np.uint16)
    return t
