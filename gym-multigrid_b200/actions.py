"""Action enums of the reference (core/agent.py:32-67); torch-free so host-side policies can import them."""
import enum


class CollectActions(enum.IntEnum):   # core/agent.py:32-36
    north = 0
    east = 1
    south = 2
    west = 3


class MazeActions(enum.IntEnum):      # core/agent.py:54-67 (CtfActions has the same members)
    stay = 0
    left = 1
    down = 2
    right = 3
    up = 4


CtfActions = MazeActions
