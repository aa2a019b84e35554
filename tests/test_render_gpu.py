"""GPU parity for `render()` (mg_render): frames recorded from the unmodified reference, and the oracle on stepped states."""
import numpy as np
import pytest
import torch

import oracle as oc
from replay import load_golden

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("stem,env_id", [("render_clustered", "multigrid-collect-respawn-clustered-v0"),
                                         ("render_rooms", "multigrid-collect-rooms-respawn-v0"),
                                         ("render_quadrants15", "multigrid-collect-quadrants15-v0")])
def test_collect_frames_match_reference(stem, env_id, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    for ts in (32, 8):
        obs, want = g[f"grid_obs_{ts}"], g[f"frames_{ts}"]
        env = mg.make_vec(env_id, len(obs), autoreset=False)
        env.set_state_from_obs(obs, g[f"pos_{ts}"])
        got = env.render(tile_size=ts)
        assert got.dtype == torch.uint8 and tuple(got.shape) == want.shape
        assert np.array_equal(_np(got), want), f"{stem} tile_size {ts}"
        assert env.status() == 0
        env.close()


def test_maze_frames_match_reference(cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden("render_maze13")
    n = len(g["pos"])
    env = mg.make_maze_vec(n, g["field_map"], autoreset=False)
    env.reset()
    a = env._agents                      # (x, y, dir, flags) per agent
    a[:, 0, 0] = torch.as_tensor(g["pos"][:, 0].astype(np.uint8), device=cuda_device)
    a[:, 0, 1] = torch.as_tensor(g["pos"][:, 1].astype(np.uint8), device=cuda_device)
    a[:, 0, 2] = torch.as_tensor(g["dir"].astype(np.uint8), device=cuda_device)
    for ts in (32, 8):
        assert np.array_equal(_np(env.render(tile_size=ts)), g[f"frames_{ts}"]), f"tile_size {ts}"
    env.close()


@pytest.mark.parametrize("ts", [32, 16, 5, 1])
def test_collect_render_vs_oracle_on_stepped_states(ts, cuda_device):
    """Philox-mode states after autoreset steps; env_ids subsets in arbitrary order, an unaligned output pointer (byte path)
    and tile sizes with and without the 16-byte fast path."""
    import gym_multigrid_b200 as mg
    n = 300
    env = mg.make_vec("multigrid-collect-rooms-respawn-v0", n, seed=3)
    env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(1)
    for _ in range(23):
        obs, *_ = env.step(torch.randint(0, 4, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8))
    want = oc.render_grid(_np(obs), ts)
    assert np.array_equal(_np(env.render(tile_size=ts)), want)
    ids = torch.tensor([299, 0, 17, 17, 150], device=cuda_device)
    assert np.array_equal(_np(env.render(env_ids=ids, tile_size=ts)), want[_np(ids)])
    buf = torch.zeros(5 * want[0].size + 1, dtype=torch.uint8, device=cuda_device)
    out = buf[1:].view(5, *want.shape[1:])          # 1 byte off 16-byte alignment
    env.render(env_ids=ids, tile_size=ts, out=out)
    assert np.array_equal(_np(out), want[_np(ids)])
    assert env.status() == 0
    env.render(env_ids=[n], tile_size=ts)           # outside [0, N): flagged, nothing out of bounds is read
    assert env.status() == 4                        # MG_ERR_OOB
    env.close()


def test_maze_render_vs_oracle_64x64(cuda_device):
    import gym_multigrid_b200 as mg
    fm = load_golden("maze_gen64")["field_map"]
    n = 40
    env = mg.make_maze_vec(n, fm, seed=5)
    env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(2)
    for _ in range(30):
        env.step(torch.randint(0, 5, (n,), generator=gen, device=cuda_device, dtype=torch.int8))
    a = _np(env._agents)[:, 0]
    for ts in (8, 3):
        assert np.array_equal(_np(env.render(tile_size=ts)), oc.render_maze(fm, a[:, :2].astype(np.int16), a[:, 2].astype(np.int8), ts))
    env.close()


@pytest.mark.parametrize("stem", ["render_ctf_2v2", "render_ctf_3v4_penalty"])
def test_ctf_frames_and_background_state_match_reference(stem, cuda_device):
    """Recorded reference episodes replayed through the CUDA step (trace mode): the sticky background colour of every agent after
    every step and every recorded frame (grey terminated agents, agents on foreign territory, both tile sizes)."""
    import gym_multigrid_b200 as mg
    from test_oracle_render_golden import replay_ctf_render
    g = load_golden(stem)
    E, T, nb = g["actions"].shape
    nr = int(g["meta_num_red"])
    env = mg.make_ctf_vec(E, g["field_map"], num_blue_agents=nb, num_red_agents=nr,
                          obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]), autoreset=False)
    env.set_trace(blue_place=g["blue_place"], red_place=g["red_place"])
    env.reset()

    def step(actions, tr):
        env.set_trace(**tr)
        env.step(torch.as_tensor(actions.astype(np.int8), device=cuda_device))

    seen = 0
    for t, eps, rows in replay_ctf_render(g, step, lambda: (_np(env.agent_pos), _np(env.agent_dir), _np(env.agent_flags))):
        for ts in (32, 8):
            got = env.render(env_ids=torch.as_tensor(eps, device=cuda_device), tile_size=ts)
            assert np.array_equal(_np(got), g[f"frames_{ts}"][rows]), f"frames after step {t}, tile_size {ts}"
        seen += len(rows)
    assert seen == len(g["frame_step"]) and env.status() == 0
    env.close()


def test_ctf_render_vs_oracle_philox(cuda_device):
    """Philox-mode CtF (2v2 with collision penalty, and 1v1 which never recolours) after autoreset steps: frames and flags."""
    import gym_multigrid_b200 as mg
    fm = load_golden("ctf_2v2")["field_map"]
    for nb, nr, v1 in ((2, 2, False), (1, 1, True)):
        n = 200
        kw = dict(max_steps=40, seed=13)
        env = mg.make_ctf1v1_vec(n, fm, **kw) if v1 else mg.make_ctf_vec(n, fm, num_blue_agents=nb, num_red_agents=nr, obstacle_penalty_ratio=0.5, **kw)
        o = (oc.CtfOracle(fm, n, 1, 1, max_steps=40, variant_1v1=1) if v1 else oc.CtfOracle(fm, n, nb, nr, obstacle_penalty_ratio=0.5, max_steps=40))
        env.reset(); o.reset(oc.map_rng(mode=1, seed=13))
        gen = torch.Generator(device=cuda_device).manual_seed(6)
        for _ in range(25):
            act = torch.randint(0, 5, (n, nb), generator=gen, device=cuda_device, dtype=torch.int8)
            env.step(act)
            o.step(_np(act), oc.map_rng(mode=1, seed=13), autoreset=True)
        assert np.array_equal(_np(env.agent_flags), o.flags) and np.array_equal(_np(env.agent_pos), o.pos)
        for ts in (32, 6):
            assert np.array_equal(_np(env.render(tile_size=ts)), oc.render_ctf(fm, o.pos, o.dir, o.flags, nb, ts, variant_1v1=v1))
        env.close()


def test_single_env_adaptors_render_like_the_reference(cuda_device):
    """`env.render()` of the reference-style classes: ndarray (H * 32, W * 32, 3) uint8 as MultiGridEnv.render returns."""
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.single_env import CtFMvNEnv, MazeSingleAgentEnv
    env = mg.make("multigrid-collect-respawn-clustered-v0")
    obs, _ = env.reset()
    f = env.render()
    assert isinstance(f, np.ndarray) and f.shape == (320, 320, 3) and f.dtype == np.uint8
    assert np.array_equal(f, oc.render_grid(obs[None], 32)[0])
    env.close()
    g = load_golden("render_maze13")
    m = MazeSingleAgentEnv(g["field_map"])
    m.reset()
    assert m.render().shape == (13 * 32, 13 * 32, 3)
    m.close()
    c = CtFMvNEnv(load_golden("ctf_2v2")["field_map"])
    c.reset()
    assert c.render().shape == (320, 320, 3)
    c.close()
    wf = mg.make_wildfire_vec(4, size=8, num_agents=2)
    assert wf.render() is None
    wf.close()
