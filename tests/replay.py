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
