"""gym-multigrid_b200: B200-native batched gridworld simulator for the hot path of
Tran-Research-Group/gym-multigrid (MultiGridEnv.step + Grid.encode), behind the reference's
gymnasium surface.

    import gym_multigrid_b200 as mg
    envs = mg.make_vec("multigrid-collect-respawn-clustered-v0", num_envs=65536, device="cuda:0")
    obs, info = envs.reset()
    obs, rewards, terminated, truncated, info = envs.step(actions)      # CUDA tensors, no host sync

    env = mg.make("multigrid-collect-respawn-clustered-v0")              # reference-style single env

The compute path is hand-written CUDA for sm_100a behind the C ABI in include/multigrid_b200.h;
there is no CPU fallback (construction raises when the library or a B200 is missing).
"""
from .registration import registry, register, register_with_gymnasium, spec  # noqa: F401
from .registration import COLLECT_CLASSES as _COLLECT_CLASSES

__version__ = "0.1.0"


def _collect_kwargs(s, overrides):
    cls = s.entry_point.split(":")[-1]
    if cls not in _COLLECT_CLASSES:
        raise NotImplementedError(f"entry point {s.entry_point!r} has no CUDA implementation")
    layout, fixed = _COLLECT_CLASSES[cls]
    kw = dict(s.kwargs)
    kw.update(layout=layout, fixed_horizon=fixed, max_episode_steps=s.max_episode_steps)
    kw.update(overrides)
    return kw


def make_vec(env_id: str, num_envs: int, device="cuda:0", seed: int = 0, autoreset: bool = True, **overrides):
    """gymnasium.vector-style constructor keyed by the reference's env ids (gym_multigrid/__init__.py:6-147)."""
    from .vector_env import CollectVecEnv
    kw = _collect_kwargs(spec(env_id), overrides)
    return CollectVecEnv(num_envs, device=device, seed=seed, autoreset=autoreset, **kw)


def make(env_id: str, device="cuda:0", seed: int = 0, **overrides):
    """gymnasium.make-style constructor: one env with the reference's reset/step signatures."""
    from .vector_env import CollectEnv
    kw = _collect_kwargs(spec(env_id), overrides)
    return CollectEnv(device=device, seed=seed, **kw)


def make_maze_vec(num_envs: int, map_path, **kwargs):
    """Batched `MazeSingleAgentEnv` (envs/maze.py); kwargs as the reference constructor (maze.py:31-40)."""
    from .map_env import MazeVecEnv
    return MazeVecEnv(num_envs, map_path, **kwargs)


def make_ctf_vec(num_envs: int, map_path, **kwargs):
    """Batched `CtFMvNEnv` (envs/ctf.py:657-1433); kwargs as the reference constructor (ctf.py:662-679)."""
    from .map_env import CtfVecEnv
    return CtfVecEnv(num_envs, map_path, **kwargs)


def make_wildfire_vec(num_envs: int, **kwargs):
    """Batched Wildfire (extension; the reference has no Wildfire code - see include/multigrid_b200.h for the rules)."""
    from .wildfire_env import WildfireVecEnv
    return WildfireVecEnv(num_envs, **kwargs)


def make_ctf1v1_vec(num_envs: int, map_path, **kwargs):
    """Batched `Ctf1v1Env` (envs/ctf.py:50-654); kwargs as the reference constructor (ctf.py:55-70)."""
    from .map_env import Ctf1v1VecEnv
    return Ctf1v1VecEnv(num_envs, map_path, **kwargs)


def make_generic_vec(num_envs: int, width: int, **kwargs):
    """Batched base-class `MultiGridEnv.step` with DefaultWorld (multigrid.py:397-483; layout injected, see generic_env.py)."""
    from .generic_env import GenericVecEnv
    return GenericVecEnv(num_envs, width, **kwargs)


_POLICIES = ("CtfPolicy", "RwPolicy", "DestinationPolicy", "FightPolicy", "CapturePolicy", "PatrolPolicy", "PatrolFightPolicy")


def __getattr__(name):
    """Reference-style single-env classes (same names as gym_multigrid.envs; imported lazily, they need torch + CUDA) and the
    host-side pieces user code imports beside them: action enums, world tables, the scripted CtF opponents."""
    if name in ("MazeSingleAgentEnv", "CtFMvNEnv", "Ctf1v1Env"):
        from . import single_env
        return getattr(single_env, name)
    if name in ("MazeActions", "CtfActions", "CollectActions"):
        from . import actions
        return getattr(actions, name)
    if name in ("DefaultWorld", "CollectWorld", "CtfWorld", "MazeWorld"):
        from . import world
        return getattr(world, name)
    if name in _POLICIES:
        from .policy.ctf import heuristic
        return getattr(heuristic, name)
    raise AttributeError(f"module 'gym_multigrid_b200' has no attribute {name!r}")
