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
