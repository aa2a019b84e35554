"""Helpers shared by the CPU (oracle) and GPU (CUDA through the C ABI) parity tests:
loading golden fixtures recorded from the reference and turning them into replay inputs."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fixture stem -> (layout, fixed_horizon) ; the reference class behind each registered id
COLLECT_FIXTURES = {
    "collect_respawn_clustered": ("quadrants_respawn", False),
    "collect_even": ("even_dist", False),
    "collect_single": ("even_dist", False),
    "collect_quadrants": ("quadrants", False),
    "collect_rooms": ("rooms", False),
    "collect_rooms_fixed": ("rooms", True),
    "collect_rooms_respawn": ("rooms", True),
    "collect_respawn": ("even_dist", False),
    "collect_quadrants15": ("quadrants", False),
    # constructor arguments outside the registry (oracle/gen_golden.py COLLECT_VARIANTS): rewards / counters that differ between
    # balls placed by _gen_grid and by _respawn, and between type index and colour index
    "collect_respawn_clustered_rewards": ("quadrants_respawn", False),
    "collect_respawn_permuted": ("even_dist", False),
}


def load_golden(stem):
    with np.load(os.path.join(GOLDEN, stem + ".npz")) as z:
        return {k: z[k] for k in z.files}


def collect_kwargs(g, stem):
    layout, fixed = COLLECT_FIXTURES[stem]
    return dict(size=int(g["meta_size"]), num_balls=int(g["meta_num_balls"]),
                agents_index=[int(v) for v in g["meta_agents_index"]],
                balls_index=[int(v) for v in g["meta_balls_index"]],
                balls_reward=[float(v) for v in g["meta_balls_reward"]],
                respawn=bool(g["meta_respawn"]), layout=layout, fixed_horizon=fixed,
                time_limit=int(g["meta_time_limit"]))


def step_inputs(g, t, tile=1):
    """Replay inputs of step t for every episode (finished episodes get no-op actions and an
    identity order so they stay frozen); optionally tiled `tile` times along the env axis."""
    E, T, A = g["actions"].shape
    live = g["length"] > t
    act = np.where(live[:, None], g["actions"][:, t], -1).astype(np.int8)
    order = np.where(live[:, None], g["order"][:, t], np.arange(A, dtype=np.uint8)[None]).astype(np.uint8)
    draws = g["draws"][:, t]
    n_draws = np.where(live, g["n_draws"][:, t], 0).astype(np.int32)
    if tile > 1:
        act, order, draws, n_draws, live = (np.concatenate([x] * tile) for x in (act, order, draws, n_draws, live))
    return act, order, draws, n_draws, live


def expected(g, t, tile=1):
    out = {k: g[k][:, t] for k in ("obs", "rewards", "terminated", "truncated", "info", "pos", "collected")}
    if tile > 1:
        out = {k: np.concatenate([v] * tile) for k, v in out.items()}
    return out


# --------------------------------------------------------------------------- CtF heuristic policies (tests/test_policies.py)
POLICY_NAMES = ("FightPolicy", "CapturePolicy", "PatrolPolicy", "PatrolFightPolicy")


def policy_maps():
    """Field maps the policy fixtures were recorded on: the reference's 10x10 board (from the CtF fixture), a 14x11 map
    with irregular territories and obstacles, and - for `a_star` alone - a 12x9 map that also holds the value 8, the
    only one the reference's A* treats as blocking (policy/ctf/utils.py:73)."""
    board = load_golden("ctf_2v2")["field_map"].astype(np.float64)
    rng = np.random.default_rng(77)
    wide = np.where(np.arange(14)[:, None] + rng.integers(-2, 3, size=(14, 11)) < 7, 1.0, 0.0)
    wide[rng.random(wide.shape) < 0.08] = 6.0
    wide[1, 2], wide[12, 8] = 5.0, 4.0
    walls = np.zeros((12, 9))
    walls[rng.random(walls.shape) < 0.28] = 8.0
    walls[3:6, 4] = 8.0
    walls[9:12, 6], walls[9, 6:9] = 8.0, 8.0        # (10..11, 7..8) is sealed off: routes into it do not exist
    return {"board": board, "wide": wide, "walls": walls}


def policy_observation(field_map, blue, red, terminated=None):
    """The dictionary `CtFMvNEnv._get_dict_obs` hands to `policy.act` (ctf.py:1112-1135) for agents at `blue` / `red`."""
    cells = lambda v: np.array(list(zip(*np.where(field_map == v)))).flatten()   # noqa: E731
    blue, red = np.asarray(blue), np.asarray(red)
    return {"blue_agent": blue.flatten(), "red_agent": red.flatten(),
            "blue_flag": np.array(list(zip(*np.where(field_map == 4)))[0]), "red_flag": np.array(list(zip(*np.where(field_map == 5)))[0]),
            "blue_territory": cells(0), "red_territory": cells(1), "obstacle": cells(6),
            "terminated_agents": np.zeros(len(blue) + len(red), np.int64) if terminated is None else np.asarray(terminated)}


def ctf_session_schedule(g):
    """Golden files of CONSECUTIVE episodes of one env instance (oracle/gen_golden.py gen_ctf_carry; episode index =
    session * K + k): every session walks through its K episodes at its own pace.  Yields
      ("reset", mask[S], ep[S])           - sessions in `mask` start episode ep[s] now (a masked reset with that episode's placements)
      ("step", live[S], ep[S], t[S])      - step t[s] of episode ep[s] for the sessions in `live` (the others have finished all
                                            their episodes: feed them anything valid and do not compare)."""
    S, K = int(g["meta_sessions"]), int(g["meta_episodes_per_session"])
    length = g["length"].reshape(S, K)
    cur_ep, cur_t, need = np.zeros(S, np.int64), np.zeros(S, np.int64), np.ones(S, bool)
    base = np.arange(S) * K
    while True:
        live = cur_ep < K
        if not live.any():
            return
        ep = base + np.minimum(cur_ep, K - 1)
        if (need & live).any():
            yield ("reset", need & live, ep)
            need[:] = False
        yield ("step", live.copy(), ep, cur_t.copy())
        cur_t[live] += 1
        done = live & (cur_t >= length[np.arange(S), np.minimum(cur_ep, K - 1)])
        cur_ep[done] += 1
        cur_t[done] = 0
        need |= done
