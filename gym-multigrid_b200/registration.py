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


def register_with_gymnasium(prefix: str = "", device: str = "cuda:0") -> list[str]:
    """When `gymnasium` is installed: make `gymnasium.make(id)` / `gymnasium.make_vec(id, num_envs=N)` resolve the reference's
    ids (optionally prefixed, e.g. "b200/") to the CUDA-backed envs of this package, with the registered kwargs and
    `max_episode_steps` of gym_multigrid/__init__.py:6-147.  The single-env entry point applies the TimeLimit itself
    (the library counts steps on the device), so no TimeLimit wrapper is requested from gymnasium.  Returns the ids registered."""
    import gymnasium  # noqa: PLC0415 - optional dependency

    def single(env_id):
        def make_single(**kwargs):
            from . import make
            return make(env_id, device=kwargs.pop("device", device), **kwargs)
        return make_single

    def vector(env_id):
        def make_vector(num_envs=1, **kwargs):
            from . import make_vec
            kwargs.pop("vectorization_mode", None)
            return make_vec(env_id, num_envs, device=kwargs.pop("device", device), **kwargs)
        return make_vector

    done = []
    for env_id in registry:
        gymnasium.register(id=prefix + env_id, entry_point=single(env_id), vector_entry_point=vector(env_id),
                           max_episode_steps=None, disable_env_checker=True, order_enforce=False)
        done.append(prefix + env_id)
    return done
