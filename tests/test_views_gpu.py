"""GPU parity for partial observations (MultiGridEnv.gen_obs) through mg_gen_obs."""
import numpy as np
import pytest
import torch

import oracle as oc
from replay import load_golden

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("stem,env_id", [("partial_rooms", "multigrid-collect-rooms-respawn-v0"),
                                         ("partial_clustered", "multigrid-collect-respawn-clustered-v0"),
                                         ("partial_quadrants15", "multigrid-collect-quadrants15-v0")])
def test_collect_views_match_reference(stem, env_id, cuda_device):
    """The reference's gen_obs_grid + encode_for_agents outputs, every recorded state, grouped by (V, see_through)."""
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    for V in (3, 5, 7):
        for st in (False, True):
            sel = np.where((g["V"] == V) & (g["see_through"] == st))[0]
            if len(sel) == 0:
                continue
            env = mg.make_vec(env_id, len(sel), autoreset=False)
            env.set_state_from_obs(g["grid_obs"][sel], g["pos"][sel])
            out = env.gen_obs(view_size=V, see_through_walls=st, dirs=g["dirs"][sel])
            assert np.array_equal(_np(out), g["views"][sel][:, :, :V, :V]), f"V={V} see_through={st}"
            env.close()


@pytest.mark.parametrize("V", [1, 3, 7, 9, 15])
def test_collect_views_match_oracle_at_scale(V, cuda_device):
    import gym_multigrid_b200 as mg
    n = 5000
    env = mg.make_vec("multigrid-collect-rooms-respawn-v0", n, seed=3)
    env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(1)
    for _ in range(7):
        env.step(torch.randint(0, 4, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8))
    dirs = torch.randint(0, 4, (n, 2), generator=gen, device=cuda_device, dtype=torch.uint8)
    for st in (False, True):
        out = env.gen_obs(view_size=V, see_through_walls=st, dirs=dirs)
        want = oc.partial_view3(_np(env.grid), _np(env.agent_pos), 10, 10, V, st, dirs=_np(dirs))
        assert np.array_equal(_np(out), want)
    out = env.gen_obs(view_size=V)   # default: every Collect agent faces up (dir 3)
    assert np.array_equal(_np(out), oc.partial_view3(_np(env.grid), _np(env.agent_pos), 10, 10, V, False))
    env.close()


@pytest.mark.parametrize("stem", ["maze_board13", "maze_gen64"])
def test_maze_views_match_composed_oracle(stem, cuda_device):
    """BASELINE config 4 (Maze + partial view): reference dynamics x the reference's gen_obs algorithm.  The oracle side
    draws each env's MazeWorld grid (Floor white / Flag red / Obstacle grey / Agent blue) and runs the pinned
    partial-view restatement on it; cells outside the map use the documented filler (3, 7, 1)."""
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    fm = g["field_map"]
    S = fm.shape[0]
    n = 777
    env = mg.make_maze_vec(n, fm, seed=5)
    env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(2)
    packed_map = np.where(fm == 0, 0 | 10 << 2, np.where(fm == 2, 2 | 0 << 2, 3 | 7 << 2)).astype(np.uint8)
    for t in range(12):
        env.step(torch.randint(0, 5, (n,), generator=gen, device=cuda_device, dtype=torch.int8))
        pos, d = _np(env.agent_pos), _np(env.agent_dir)
        grids = np.repeat(packed_map.reshape(1, -1), n, axis=0)
        grids[np.arange(n), pos[:, 0, 0].astype(int) * S + pos[:, 0, 1]] = (1 | 4 << 2) | (d[:, 0] << 6)
        for V, st in ((7, False), (5, True)):
            out = env.gen_obs(view_size=V, see_through_walls=st)
            want = oc.partial_view3(grids, pos, S, S, V, st, dirs=d, oob_code=3 | 7 << 2 | 1 << 6, opaque_rule=1)
            assert np.array_equal(_np(out), want), f"step {t} V={V}"
    env.close()


@pytest.mark.parametrize("table", ["1", "0"])
@pytest.mark.parametrize("V,st", [(7, False), (5, True), (3, False)])
def test_maze_fused_step_and_partial_view(V, st, table, cuda_device, monkeypatch):
    """observation_option="partial": the step kernel itself writes the gen_obs views (mg_set_partial_obs) - computed per step (default), or copied from
    the memoised table of all S*S*4 agent states (MG_VIEW_TABLE=1, what maps too large for shared memory use); they must equal mg_gen_obs on
    the post-step state and the composed oracle, including across same-step autoresets and a ragged tile."""
    import gym_multigrid_b200 as mg
    monkeypatch.setenv("MG_VIEW_TABLE", table)
    g = load_golden("maze_gen64")
    fm = g["field_map"]
    S, n = fm.shape[0], 1000
    env = mg.make_maze_vec(n, fm, seed=6, max_steps=9, observation_option="partial", view_size=V, see_through_walls=st)
    packed_map = np.where(fm == 0, 0 | 10 << 2, np.where(fm == 2, 2 | 0 << 2, 3 | 7 << 2)).astype(np.uint8)

    def want():
        pos, d = _np(env.agent_pos), _np(env.agent_dir)
        grids = np.repeat(packed_map.reshape(1, -1), n, axis=0)
        grids[np.arange(n), pos[:, 0, 0].astype(int) * S + pos[:, 0, 1]] = (1 | 4 << 2) | (d[:, 0] << 6)
        return oc.partial_view3(grids, pos, S, S, V, st, dirs=d, oob_code=3 | 7 << 2 | 1 << 6, opaque_rule=1)

    obs, _ = env.reset()
    assert tuple(obs.shape) == (n, 1, V, V, 3) and np.array_equal(_np(obs), want())
    gen = torch.Generator(device=cuda_device).manual_seed(3)
    for t in range(25):
        obs, rew, term, trunc, _ = env.step(torch.randint(0, 5, (n,), generator=gen, device=cuda_device, dtype=torch.int8))
        assert np.array_equal(_np(obs), _np(env.gen_obs(V, st))), f"step {t}: fused views != mg_gen_obs"
        assert np.array_equal(_np(obs), want()), f"step {t}: fused views != oracle"
    assert int(env.episode_count.min()) >= 3
    host_obs = env.step(np.zeros(n, np.int8))[0]
    assert host_obs.shape == (n, 1, V, V, 3) and np.array_equal(host_obs, want())
    env.set_partial_obs(0)
    obs, *_ = env.step(torch.zeros(n, dtype=torch.int8, device=cuda_device))
    assert tuple(obs.shape) == (n, S, S)
    env.close()


@pytest.mark.parametrize("stem,env_id", [("toroid_clustered", "multigrid-collect-respawn-clustered-v0"),
                                         ("toroid_rooms", "multigrid-collect-rooms-respawn-v0"),
                                         ("toroid_single", "multigrid-collect-single-v0")])
def test_toroid_matches_reference(stem, env_id, cuda_device):
    """The reference's ToroidObservation.observation outputs on recorded states (incl. single-agent depth 4)."""
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.wrappers import ToroidObservation
    g = load_golden(stem)
    n = len(g["pos"])
    env = ToroidObservation(mg.make_vec(env_id, n, autoreset=False))
    env.env.set_state_from_obs(g["grid_obs"], g["pos"])
    out = env.observation()
    assert out.dtype == torch.float32 and tuple(out.shape) == g["toroid"].shape
    assert np.array_equal(_np(out), g["toroid"])
    env.close()


def test_toroid_wrapper_matches_oracle_while_stepping(cuda_device):
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.wrappers import ToroidObservation
    n = 3000
    env = ToroidObservation(mg.make_vec("multigrid-collect-rooms-respawn-v0", n, seed=8))
    obs, _ = env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(4)
    for t in range(60):
        obs, rew, term, trunc, _ = env.step(torch.randint(0, 4, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8))
        if t % 10 == 9:
            assert np.array_equal(_np(obs), oc.toroid(_np(env.grid), _np(env.agent_pos), 10, 3))
    env.close()
