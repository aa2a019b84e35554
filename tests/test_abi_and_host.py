"""CPU suite: the C-ABI library loads and exports every symbol include/multigrid_b200.h
declares (no compute calls without a GPU), and the host-side mirror of the reference's
registration / spaces is faithful."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "multigrid_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gym_multigrid_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/multigrid_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert lib.mg_abi_version() == 1


def test_config_struct_matches_header_layout():
    from gym_multigrid_b200 import _lib
    # uint32 + int32 + 2*int64 + 4*int32 + 8*int32 + 8*int32 + 8*double + 7*int32 (+pad) + uint64
    assert C.sizeof(_lib.Config) == 4 + 4 + 16 + 16 + 32 + 32 + 64 + 28 + 4 + 8
    assert C.sizeof(_lib.StepIO) == 6 * 8
    assert C.sizeof(_lib.MapConfig) == 120 and C.sizeof(_lib.MapTrace) == 8 * 8
    assert C.sizeof(_lib.WildfireConfig) == 4 + 4 + 16 + 12 + 128 + 4 + 20 + 4 + 8 + 8
    assert C.sizeof(_lib.Trace) == 9 * 8
    assert C.sizeof(_lib.RedPolicies) == 4 + 4 + 64 + 128 + 4 * 8 + 4 + 4


def test_struct_sizes_agree_with_the_c_compiler(tmp_path):
    """The ctypes mirrors against `sizeof` as gcc sees include/multigrid_b200.h (every struct the ABI checks by struct_size)."""
    import shutil
    import subprocess
    from gym_multigrid_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    names = {"mg_config": _lib.Config, "mg_map_config": _lib.MapConfig, "mg_wildfire_config": _lib.WildfireConfig,
             "mg_generic_config": _lib.GenericConfig, "mg_red_policies": _lib.RedPolicies, "mg_step_io": _lib.StepIO, "mg_trace": _lib.Trace, "mg_map_trace": _lib.MapTrace}
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "multigrid_b200.h"\nint main(void) {\n'
                   + "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n, cls in names.items():
        assert int(out[n]) == C.sizeof(cls), n


def test_create_fails_loudly_without_gpu_or_with_bad_abi():
    import torch
    from gym_multigrid_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config()
    h = C.c_void_p()
    cfg.struct_size = 3
    assert lib.mg_create(C.byref(cfg), 0, C.byref(h)) != 0 and "size mismatch" in _lib.last_error(None)
    if not torch.cuda.is_available():
        import gym_multigrid_b200 as mg
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            mg.make_vec("multigrid-collect-respawn-clustered-v0", 8)


def test_registry_mirrors_reference_ids_and_kwargs():
    import gym_multigrid_b200 as mg
    ids = {
        "multigrid-collect-v0": ("CollectGameEvenDist", 100, [3, 5], False, 10, 15),
        "multigrid-collect-single-v0": ("CollectGameEvenDist", 100, [3], False, 10, 15),
        "multigrid-collect-quadrants-v0": ("CollectGameQuadrants", 100, [3, 5], False, 10, 15),
        "multigrid-collect-rooms-v0": ("CollectGameRooms", 100, [3, 5], False, 10, 15),
        "multigrid-collect-rooms-fixed-horizon-v0": ("CollectGameRoomsFixedHorizon", 100, [3, 5], False, 10, 15),
        "multigrid-collect-rooms-respawn-v0": ("CollectGameRoomsFixedHorizon", 50, [3, 5], True, 10, 15),
        "multigrid-collect-respawn-v0": ("CollectGameEvenDist", 50, [3, 5], True, 10, 15),
        "multigrid-collect-respawn-clustered-v0": ("CollectGameQuadrantsRespawn", 50, [3, 5], True, 10, 15),
        "multigrid-collect-quadrants15-v0": ("CollectGameQuadrants", None, [3, 5], False, 15, 30),
    }
    assert set(mg.registry) == set(ids)
    for k, (cls, steps, agents, respawn, size, balls) in ids.items():
        s = mg.spec(k)
        assert s.entry_point == f"gym_multigrid.envs:{cls}" and s.max_episode_steps == steps
        assert s.kwargs == dict(size=size, num_balls=balls, agents_index=agents, balls_index=[0, 1, 2],
                                balls_reward=[1, 1, 1], respawn=respawn)
    assert mg.spec("gym_multigrid:multigrid-collect-v0").id == "multigrid-collect-v0"
    with pytest.raises(KeyError):
        mg.spec("multigrid-nope-v0")


def test_golden_registry_agrees_with_fixture_metadata():
    """The ids/kwargs in our registry are the ones the reference registered when the fixtures were recorded."""
    import gym_multigrid_b200 as mg
    from replay import COLLECT_FIXTURES, load_golden
    for stem in COLLECT_FIXTURES:
        g = load_golden(stem)
        s = mg.spec(str(g["meta_env_id"]))
        assert s.kwargs["size"] == int(g["meta_size"]) and s.kwargs["num_balls"] == int(g["meta_num_balls"])
        assert s.kwargs["agents_index"] == g["meta_agents_index"].tolist()
        assert (s.max_episode_steps or 0) == int(g["meta_time_limit"])
        assert s.kwargs["respawn"] == bool(g["meta_respawn"])


def test_spaces():
    from gym_multigrid_b200.spaces import Box, Discrete, MultiDiscrete
    d = Discrete(4)
    assert d.n == 4
    b = Box(0, 255, (10, 10, 3), np.uint8)
    assert b.shape == (10, 10, 3) and b.dtype == np.uint8
    m = MultiDiscrete([5, 5])
    assert m.nvec.tolist() == [5, 5]


def test_register_with_gymnasium_hands_over_every_id(monkeypatch):
    """gymnasium is not in the image: a stand-in module records what the helper registers."""
    import sys
    import types
    import gym_multigrid_b200 as mg
    calls = []
    fake = types.ModuleType("gymnasium")
    fake.register = lambda **kw: calls.append(kw)
    monkeypatch.setitem(sys.modules, "gymnasium", fake)
    ids = mg.register_with_gymnasium(prefix="b200/")
    assert ids == ["b200/" + k for k in mg.registry] and len(calls) == len(mg.registry) == 9
    for kw in calls:
        assert callable(kw["entry_point"]) and callable(kw["vector_entry_point"]) and kw["max_episode_steps"] is None


def test_vector_env_surface_is_shared_by_every_batched_class():
    from gym_multigrid_b200.generic_env import GenericVecEnv
    from gym_multigrid_b200.map_env import CtfVecEnv, MazeVecEnv
    from gym_multigrid_b200.vector_base import VectorEnvSurface
    from gym_multigrid_b200.vector_env import CollectVecEnv
    from gym_multigrid_b200.wildfire_env import WildfireVecEnv
    for cls in (CollectVecEnv, MazeVecEnv, CtfVecEnv, WildfireVecEnv, GenericVecEnv):
        assert issubclass(cls, VectorEnvSurface)
        for name in ("reset", "step", "close", "render", "get_attr", "unwrapped", "metadata"):
            assert hasattr(cls, name), (cls.__name__, name)


def test_reference_import_paths_resolve():
    """Code written against the reference changes the package name only: the module paths its tests and script import from
    (tests/test_ctf.py:7-16, tests/test_maze.py:3, scripts/main_mvn_ctf_rl.py:7) exist under gym_multigrid_b200."""
    from gym_multigrid_b200.core.agent import CollectActions, CtfActions, MazeActions
    from gym_multigrid_b200.core.world import CollectWorld, CtfWorld, DefaultWorld, MazeWorld
    from gym_multigrid_b200.envs.collect_game import CollectEnv, CollectVecEnv
    from gym_multigrid_b200.envs.ctf import Ctf1v1Env, CtFMvNEnv
    from gym_multigrid_b200.envs.maze import MazeSingleAgentEnv
    from gym_multigrid_b200.policy.ctf.heuristic import CapturePolicy, FightPolicy, PatrolFightPolicy, PatrolPolicy, RwPolicy
    from gym_multigrid_b200.utils.map import load_text_map
    from gym_multigrid_b200.utils.misc import set_seed
    from gym_multigrid_b200.wrappers import ToroidObservation
    import gym_multigrid_b200 as mg
    assert mg.CtFMvNEnv is CtFMvNEnv and mg.Ctf1v1Env is Ctf1v1Env and mg.MazeSingleAgentEnv is MazeSingleAgentEnv
    assert mg.FightPolicy is FightPolicy and mg.CtfWorld is CtfWorld and mg.CtfActions is CtfActions is MazeActions
    assert [a.name for a in CollectActions] == ["north", "east", "south", "west"]                       # core/agent.py:32-36
    assert CollectWorld.OBJECT_TO_IDX == dict(empty=0, wall=1, ball=2, agent=3) and DefaultWorld.encode_dim == 6
    assert MazeWorld.IDX_TO_OBJECT[3] == "obstacle" and all(c is not None for c in (CollectEnv, CollectVecEnv, ToroidObservation))
    assert all(callable(f) for f in (CapturePolicy, PatrolFightPolicy, PatrolPolicy, RwPolicy, load_text_map, set_seed))
    set_seed(5)
    a = np.random.rand()
    set_seed(5)
    assert a == np.random.rand()


def test_map_helpers(tmp_path):
    """utils/map.py:7-61 - against the reference's functions where /root/reference exists, known answers everywhere."""
    from gym_multigrid_b200.utils import map as M
    p = tmp_path / "m.txt"
    p.write_text("0 1 2\n3 4 5\n")
    fm = M.load_text_map(str(p))
    assert fm.shape == (3, 2) and fm[2, 0] == 2 and fm[0, 1] == 3 and M.load_text_map(fm) is fm       # field_map[x, y] = row y, column x
    assert M.distance_points((0, 0), (3, 4)) == 5.0 and M.distance_points((0, 0), (3, 4), True) == float("inf")
    assert M.distance_area_point((0, 0), [(5, 5), (1, 1), (2, 0)]) == np.sqrt(2.0)
    assert M.closest_area_pos((0, 0), [(5, 5), (1, 1), (1, 1)]) == (1, 1) and M.position_in_positions((1, 1), [(0, 1), (1, 1)])
    import ref_harness as rh
    if not rh.reference_available():
        return
    rh.import_reference()
    from gym_multigrid.utils import map as R
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = tuple(rng.integers(0, 70, 2)), tuple(rng.integers(0, 70, 2))
        area = [tuple(v) for v in rng.integers(0, 70, (int(rng.integers(1, 9)), 2))]
        assert M.distance_points(a, b) == R.distance_points(a, b)
        assert M.distance_area_point(a, area) == R.distance_area_point(a, area)
        assert tuple(M.closest_area_pos(a, area)) == tuple(R.closest_area_pos(a, area))
        assert M.position_in_positions(a, area) == R.position_in_positions(a, area)
    assert np.array_equal(M.load_text_map(str(p)), R.load_text_map(str(p)))


def test_constants_module():
    """core/constants.py:5-74 and the colour maps of the worlds: against the oracle's rasteriser (pinned to frames recorded from the
    reference: the centre pixel of a ball tile is the ball's colour, a wall tile is grey) and, where it exists, the reference."""
    import oracle as oc
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.core import constants as K
    assert list(K.COLOR_TO_IDX) == ["red", "orange", "yellow", "green", "blue", "purple", "brown", "grey", "light_red", "light_blue"]
    assert mg.CtfWorld.COLOR_TO_IDX["blue_grey"] == 12 and mg.MazeWorld.COLOR_TO_IDX["white"] == 10 and K.TILE_PIXELS == 32
    assert [v.tolist() for v in K.DIR_TO_VEC] == [[1, 0], [0, 1], [-1, 0], [0, -1]]
    obs = np.zeros((1, 10, 1, 3), np.uint8)
    for c in range(10):
        obs[0, c, 0] = (mg.CollectWorld.OBJECT_TO_IDX["ball"], c, 0)
    frame = oc.render_grid(obs, 32)[0]
    for name, c in K.COLOR_TO_IDX.items():
        assert frame[16, 32 * c + 16].tolist() == K.COLORS[name].tolist(), name
    obs[0, 0, 0] = (mg.CollectWorld.OBJECT_TO_IDX["wall"], K.COLOR_TO_IDX["grey"], 0)
    assert oc.render_grid(obs, 32)[0][16, 16].tolist() == K.COLORS["grey"].tolist()
    import ref_harness as rh
    if not rh.reference_available():
        return
    rh.import_reference()
    from gym_multigrid.core import constants as R, world as RW
    for n in ("COLORS", "CTF_COLORS", "MAZE_COLORS"):
        a, b = getattr(R, n), getattr(K, n)
        assert list(a) == list(b) and all(np.array_equal(a[k], b[k]) for k in a), n
    assert R.COLOR_TO_IDX == K.COLOR_TO_IDX and R.STATE_TO_IDX == K.STATE_TO_IDX and R.COLOR_NAMES == K.COLOR_NAMES
    assert all(np.array_equal(x, y) for x, y in zip(R.DIR_TO_VEC, K.DIR_TO_VEC))
    for n in ("DefaultWorld", "CollectWorld", "CtfWorld", "MazeWorld"):
        a, b = getattr(RW, n), getattr(mg, n)
        assert dict(a.OBJECT_TO_IDX) == dict(b.OBJECT_TO_IDX) and dict(a.COLOR_TO_IDX) == dict(b.COLOR_TO_IDX)
        assert dict(a.IDX_TO_OBJECT) == dict(b.IDX_TO_OBJECT) and dict(a.IDX_TO_COLOR) == dict(b.IDX_TO_COLOR) and a.encode_dim == b.encode_dim


def test_reference_raises_on_non_square_maps_too():
    """`mg_create_map` refuses non-square maps.  So does the reference, by accident: it reads `height, width = field_map.shape`
    (maze.py:68-70, ctf.py:745-747) and then addresses the grid as (x < width, y < height) with x running over shape[0]
    (maze.py:184-186 -> Grid.set, grid.py:61-64), so any map with shape[0] != shape[1] trips the bounds assert while the grid
    is built - in both orientations, for Maze and CtF.  Checked live where /root/reference exists."""
    import tempfile

    import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("needs /root/reference (build container only)")
    rh.import_reference()
    from gym_multigrid.envs.ctf import CtFMvNEnv
    from gym_multigrid.envs.maze import MazeSingleAgentEnv
    from gym_multigrid.policy.ctf.heuristic import RwPolicy

    def build(m, cls, **kw):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as tmp:
            np.savetxt(tmp.name, m, fmt="%d")
        try:
            env = cls(tmp.name, **kw)
            env.reset(seed=0)
        finally:
            os.unlink(tmp.name)

    for shape in ((6, 9), (9, 6)):
        m = np.zeros(shape, int)
        m[1, 2], m[3, 3] = 2, 3
        with pytest.raises(AssertionError):
            build(m, MazeSingleAgentEnv)
        c = np.zeros(shape, int)
        c[:, shape[1] // 2:] = 1
        c[1, 1], c[2, shape[1] - 2], c[3, 2] = 4, 5, 6
        with pytest.raises(AssertionError):
            build(c, CtFMvNEnv, num_blue_agents=2, num_red_agents=2, enemy_policies=RwPolicy())
    sq = np.zeros((7, 7), int)
    sq[1, 2], sq[3, 3] = 2, 3
    build(sq, MazeSingleAgentEnv)        # the square control builds and resets
