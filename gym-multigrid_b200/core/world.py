"""`gym_multigrid.core.world` index tables (core/world.py:33-91)."""
from ..world import CollectWorld, CtfWorld, DefaultWorld, MazeWorld, World  # noqa: F401
