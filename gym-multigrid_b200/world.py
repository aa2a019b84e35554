"""Object-type index tables of the reference's worlds (core/world.py:33-91) for host-side callers (policies, user code
that reads "map" observations).  The kernels bake the same numbers in (csrc/*_params.cuh); this module is data only."""
from __future__ import annotations

from types import MappingProxyType


class World:
    def __init__(self, encode_dim: int, **object_to_idx: int):
        self.encode_dim, self.normalize_obs = encode_dim, 1
        self.OBJECT_TO_IDX = MappingProxyType(dict(object_to_idx))
        self.IDX_TO_OBJECT = MappingProxyType({v: k for k, v in object_to_idx.items()})


DefaultWorld = World(6, unseen=0, empty=1, wall=2, floor=3, door=4, key=5, ball=6, box=7, goal=8, lava=9, agent=10,
                     objgoal=11, switch=12)                                                           # world.py:33-52
CollectWorld = World(3, empty=0, wall=1, ball=2, agent=3)                                             # world.py:54-64
CtfWorld = World(3, blue_territory=0, red_territory=1, blue_agent=2, red_agent=3, blue_flag=4, red_flag=5,
                 obstacle=6)                                                                          # world.py:66-79
MazeWorld = World(3, background=0, agent=1, flag=2, obstacle=3)                                       # world.py:81-91
