"""CPU suite: the C restatement of the reference's rgb_array render (oracle/mg_oracle_render.c) against frames recorded from
the unmodified reference (`MultiGridEnv.render()`; oracle/gen_golden.py render -> tests/golden/render_*.npz)."""
import numpy as np
import pytest

import oracle as oc
from replay import load_golden

COLLECT = ["render_clustered", "render_rooms", "render_quadrants15"]


@pytest.mark.parametrize("stem", COLLECT)
@pytest.mark.parametrize("ts", [32, 8])
def test_collect_frames_match_reference(stem, ts):
    g = load_golden(stem)
    want = g[f"frames_{ts}"]
    got = oc.render_grid(g[f"grid_obs_{ts}"], ts)
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want)
    assert len({tuple(d) for d in g[f"grid_obs_{ts}"][..., 2][g[f"grid_obs_{ts}"][..., 0] == 3].reshape(-1, 1)}) >= 2   # several agent rotations drawn


@pytest.mark.parametrize("ts", [32, 8])
def test_maze_frames_match_reference(ts):
    g = load_golden("render_maze13")
    assert sorted(set(g["dir"].tolist())) == [0, 1, 2, 3]
    assert np.array_equal(oc.render_maze(g["field_map"], g["pos"], g["dir"], ts), g[f"frames_{ts}"])


def test_tile_known_answers():
    """Hand-derived from the code: a None cell is black under the two grid lines (grid.py:160-161: x or y <= 0.031 covers
    exactly the first pixel row / column at tile_size 32); a wall is its colour everywhere (object.py:181-182, COLORS grey)."""
    obs = np.zeros((1, 2, 1, 3), np.uint8)
    obs[0, 1, 0] = (1, 7, 0)     # wall, grey
    f = oc.render_grid(obs, 32)[0]
    assert f.shape == (32, 64, 3)
    empty, wall = f[:, :32], f[:, 32:]
    assert (empty[0] == 100).all() and (empty[:, 0] == 100).all() and (empty[1:, 1:] == 0).all()
    assert (wall == 100).all()
    ball = np.zeros((1, 1, 1, 3), np.uint8)
    ball[0, 0, 0] = (2, 0, 0)    # ball, red: centre pixel inside the r = 0.31 circle, corner outside
    b = oc.render_grid(ball, 32)[0]
    assert tuple(b[16, 16]) == (228, 3, 3) and tuple(b[31, 31]) == (0, 0, 0)


def replay_ctf_render(g, step_fn, state_fn):
    """Drive `step_fn(actions, trace kwargs)` over the recorded episodes; after the reset and every recorded frame step yield
    (episode indices, frame rows) so the caller can compare frames.  Shared by the oracle test here and the GPU test."""
    E, T, nb = g["actions"].shape
    n = nb + int(g["meta_num_red"])
    ident = np.arange(n, dtype=np.uint8)[None]
    fe, fs = g["frame_episode"], g["frame_step"]
    yield -1, fe[fs == -1], np.where(fs == -1)[0]
    for t in range(T):
        live = g["length"] > t
        step_fn(np.where(live[:, None], g["actions"][:, t], 0),
                dict(red_actions=g["red_actions"][:, t], order=np.where(live[:, None], g["order"][:, t], ident), blue_win=g["blue_win"][:, t]))
        pos, dirs, flags = state_fn()
        assert np.array_equal(pos[live], g["pos"][live, t]) and np.array_equal(dirs[live], g["dir"][live, t])
        bgs = (flags >> 2) & 3
        bg = np.where(bgs == 0, g["init_bg"], bgs)          # 0 = as constructed
        assert np.array_equal(bg[live], g["bg"][live, t]), f"step {t}: sticky background colour (agent.py:197-200)"
        assert np.array_equal((flags & 1)[live], g["dead"][live, t])
        rows = np.where(fs == t)[0]
        if len(rows):
            yield t, fe[rows], rows


@pytest.mark.parametrize("stem", ["render_ctf_2v2", "render_ctf_3v4_penalty"])
def test_ctf_frames_and_background_state_match_reference(stem):
    g = load_golden(stem)
    E, T, nb = g["actions"].shape
    nr = int(g["meta_num_red"])
    o = oc.CtfOracle(g["field_map"], E, nb, nr, obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]))
    obs = o.reset(oc.map_rng(mode=0, blue_place=g["blue_place"], red_place=g["red_place"]))
    assert np.array_equal(obs, g["init_obs"]) and np.array_equal(g["init_bg"], np.repeat([[1] * nb + [2] * nr], E, 0))
    seen = 0
    for t, eps, rows in replay_ctf_render(g, lambda a, tr: o.step(a, oc.map_rng(mode=0, **tr)), lambda: (o.pos, o.dir, o.flags)):
        for ts in (32, 8):
            got = oc.render_ctf(g["field_map"], o.pos[eps], o.dir[eps], o.flags[eps], nb, ts)
            assert np.array_equal(got, g[f"frames_{ts}"][rows]), f"frames after step {t}, tile_size {ts}"
        seen += len(rows)
    assert seen == len(g["frame_step"]) and o.status.value == 0
