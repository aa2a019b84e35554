"""`gym_multigrid.core.agent` action enums (core/agent.py:32-67)."""
from ..actions import CollectActions, CtfActions, MazeActions  # noqa: F401
