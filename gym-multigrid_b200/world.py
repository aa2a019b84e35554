"""Object-type index tables of the reference's worlds (core/world.py:33-91) for host-side callers (policies, user code
that reads "map" observations).  The kernels bake the same numbers in (csrc/*_params.cuh); this module is data only."""
from __future__ import annotations

from types import MappingProxyType

from .core.constants import COLORS, CTF_COLORS, MAZE_COLORS


class World:
    def __init__(self, encode_dim: int, colors=None, **object_to_idx: int):
        self.encode_dim, self.normalize_obs = encode_dim, 1
        self.OBJECT_TO_IDX = MappingProxyType(dict(object_to_idx))
        self.IDX_TO_OBJECT = MappingProxyType({v: k for k, v in object_to_idx.items()})
        if colors is not None:      # world.py:21-27: colour indices follow the palette's order
            self.COLORS = colors
            self.COLOR_TO_IDX = MappingProxyType({name: i for i, name in enumerate(colors)})
            self.IDX_TO_COLOR = MappingProxyType({i: name for i, name in enumerate(colors)})


DefaultWorld = World(6, COLORS, unseen=0, empty=1, wall=2, floor=3, door=4, key=5, ball=6, box=7, goal=8, lava=9, agent=10,
                     objgoal=11, switch=12)                                                           # world.py:33-52
CollectWorld = World(3, COLORS, empty=0, wall=1, ball=2, agent=3)                                             # world.py:54-64
CtfWorld = World(3, CTF_COLORS, blue_territory=0, red_territory=1, blue_agent=2, red_agent=3, blue_flag=4, red_flag=5,
                 obstacle=6)                                                                          # world.py:66-79
MazeWorld = World(3, MAZE_COLORS, background=0, agent=1, flag=2, obstacle=3)                                       # world.py:81-91
