"""Environment ids and constructor kwargs, identical to the reference's registrations
(gym_multigrid/__init__.py:6-147) so `make(id)` / `make_vec(id)` are drop-ins.

Each entry: reference entry-point class -> (layout, fixed_horizon) understood by the CUDA
library, the registered kwargs verbatim, and `max_episode_steps` (gymnasium TimeLimit).
"""
from __future__ import annotations

from dataclasses import dataclass, field

# reference class -> (layout, fixed_horizon)
COLLECT_CLASSES = {
    "CollectGameEvenDist": ("even_dist", False),               # collect_game.py:227-259
    "CollectGameQuadrants": ("quadrants", False),              # collect_game.py:261-300
    "CollectGameRooms": ("rooms", False),                      # collect_game.py:302-362
    "CollectGameRoomsFixedHorizon": ("rooms", True),           # collect_game.py:364-370
    "CollectGameQuadrantsRespawn": ("quadrants_respawn", False),  # collect_game.py:372-409
}


@dataclass(frozen=True)
class EnvSpec:
    id: str
    entry_point: str
    max_episode_steps: int | None
    kwargs: dict = field(default_factory=dict)


def _collect(id, cls, steps, size=10, num_balls=15, agents=(3, 5), respawn=False):
    return EnvSpec(id, f"gym_multigrid.envs:{cls}", steps, dict(
        size=size, num_balls=num_balls, agents_index=list(agents), balls_index=[0, 1, 2],
        balls_reward=[1, 1, 1], respawn=respawn))


registry: dict[str, EnvSpec] = {s.id: s for s in [
    _collect("multigrid-collect-v0", "CollectGameEvenDist", 100),                                  # :6-18
    _collect("multigrid-collect-single-v0", "CollectGameEvenDist", 100, agents=(3,)),              # :22-34
    _collect("multigrid-collect-quadrants-v0", "CollectGameQuadrants", 100),                       # :38-50
    _collect("multigrid-collect-rooms-v0", "CollectGameRooms", 100),                               # :54-66
    _collect("multigrid-collect-rooms-fixed-horizon-v0", "CollectGameRoomsFixedHorizon", 100),     # :71-83
    _collect("multigrid-collect-rooms-respawn-v0", "CollectGameRoomsFixedHorizon", 50, respawn=True),   # :88-100
    _collect("multigrid-collect-respawn-v0", "CollectGameEvenDist", 50, respawn=True),             # :105-117
    _collect("multigrid-collect-respawn-clustered-v0", "CollectGameQuadrantsRespawn", 50, respawn=True),  # :122-134
    _collect("multigrid-collect-quadrants15-v0", "CollectGameQuadrants", None, size=15, num_balls=30),    # :136-147
]}


def spec(env_id: str) -> EnvSpec:
    key = env_id.split(":")[-1]  # accepts "gym_multigrid:multigrid-collect-v0" as tests/test_collect.py:12 does
    if key not in registry:
        raise KeyError(f"unknown env id {env_id!r}; known: {sorted(registry)}")
    return registry[key]


def register(id: str, entry_point: str, max_episode_steps: int | None = None, kwargs: dict | None = None):
    """Same signature as gymnasium.register for the subset the reference uses."""
    registry[id] = EnvSpec(id, entry_point, max_episode_steps, dict(kwargs or {}))
