"""`from gym_multigrid_b200.envs.maze import MazeSingleAgentEnv` (reference: envs/maze.py:26-377; tests/test_maze.py:3)."""
from ..map_env import MazeVecEnv  # noqa: F401
from ..single_env import MazeSingleAgentEnv  # noqa: F401
