"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz (deterministic)

The committed .npz files are what the CPU and GPU parity tests replay; this script is the
provenance record the task statement asks for ("commit the vectors ... together with the
script that made them").  Every array is produced by reference code executing under
oracle/ref_harness.py's shim with RNG taps; nothing is synthesised here except the
action streams (seeded numpy Generator, including a few out-of-range actions, and a
greedy ball-seeking policy on some episodes so that `terminated` is exercised).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# id -> (file stem, episodes, greedy fraction)
COLLECT_PLAN = {
    "multigrid-collect-respawn-clustered-v0": ("collect_respawn_clustered", 256, 0.25),
    "multigrid-collect-v0": ("collect_even", 12, 0.5),
    "multigrid-collect-single-v0": ("collect_single", 8, 0.5),
    "multigrid-collect-quadrants-v0": ("collect_quadrants", 12, 0.5),
    "multigrid-collect-rooms-v0": ("collect_rooms", 12, 0.5),
    "multigrid-collect-rooms-fixed-horizon-v0": ("collect_rooms_fixed", 8, 0.5),
    "multigrid-collect-rooms-respawn-v0": ("collect_rooms_respawn", 12, 0.25),
    "multigrid-collect-respawn-v0": ("collect_respawn", 12, 0.25),
    "multigrid-collect-quadrants15-v0": ("collect_quadrants15", 6, 0.5),
}
# Registered ids with constructor arguments the registry never uses, chosen to tell apart what coincides under balls_index =
# [0, 1, 2] / balls_reward = [1, 1, 1]: a ball's reward is an attribute of the ball OBJECT - balls_reward[type] (or the literal 1 of
# QuadrantsRespawn, collect_game.py:393) for balls placed by _gen_grid, balls_reward[COLOUR index] for balls placed by _respawn
# (:130, :409) - and the info counter is indexed by the colour index (:147).   stem -> (id, overrides, episodes, greedy fraction)
COLLECT_VARIANTS = {
    "collect_respawn_clustered_rewards": ("multigrid-collect-respawn-clustered-v0", dict(balls_reward=[2.0, 3.0, 5.0]), 24, 0.75),
    "collect_respawn_permuted": ("multigrid-collect-respawn-v0", dict(balls_index=[2, 0, 1], balls_reward=[1.0, 2.0, 4.0]), 24, 0.75),
}


class GreedyActions:
    """Action source handed to record_collect_episode: mostly walks towards the nearest ball."""

    def __init__(self, env, seed, eps=0.2):
        self.env, self.rng, self.eps = env, np.random.default_rng(seed), eps

    def random(self, n):
        return self.rng.random(n)

    def choice(self, a, size):
        return self.rng.choice(a, size=size)

    def integers(self, lo, hi, size):
        env = self.env
        acts = self.rng.integers(lo, hi, size=size)
        balls = [(i, j) for i in range(env.width) for j in range(env.height)
                 if (o := env.grid.get(i, j)) is not None and o.type == "ball"]
        if not balls:
            return acts
        for k, a in enumerate(env.agents):
            if self.rng.random() < self.eps:
                continue
            x, y = int(a.pos[0]), int(a.pos[1])
            bx, by = min(balls, key=lambda b: abs(b[0] - x) + abs(b[1] - y))
            if abs(bx - x) >= abs(by - y) and bx != x:
                acts[k] = 1 if bx > x else 3
            elif by != y:
                acts[k] = 2 if by > y else 0
        return acts


def gen_collect(only_variants=False):
    rh.import_reference()
    plan = [] if only_variants else [(env_id, {}, stem, episodes, gf) for env_id, (stem, episodes, gf) in COLLECT_PLAN.items()]
    plan += [(env_id, ov, stem, episodes, gf) for stem, (env_id, ov, episodes, gf) in COLLECT_VARIANTS.items()]
    for env_id, overrides, stem, episodes, greedy_frac in plan:
        env, time_limit = rh.make_collect(env_id, **overrides)
        eps = []
        for seed in range(episodes):
            greedy = seed < int(round(greedy_frac * episodes))
            src = GreedyActions(env, 1000 + seed) if greedy else np.random.default_rng(1000 + seed)
            eps.append(rh.record_collect_episode(env, time_limit, seed, src, oob_action_prob=0.03))
        T = max(e["length"] for e in eps)
        K = max(2, max(int(e["n_draws"].max()) for e in eps))
        packed = rh.pack_collect_episodes(eps, T, K)
        from gymnasium.envs.registration import registry
        kw = dict(registry[env_id]["kwargs"])
        kw.update(overrides)
        packed["meta_env_id"] = np.array(env_id)
        packed["meta_time_limit"] = np.array(time_limit or 0)
        packed["meta_size"] = np.array(kw["size"])
        packed["meta_num_balls"] = np.array(kw["num_balls"])
        packed["meta_agents_index"] = np.array(kw["agents_index"])
        packed["meta_balls_index"] = np.array(kw["balls_index"])
        packed["meta_balls_reward"] = np.array(kw["balls_reward"], dtype=np.float64)
        packed["meta_respawn"] = np.array(kw["respawn"])
        path = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path, **packed)
        term = int(packed["terminated"].any(axis=1).sum())
        lost = 0
        print(f"{stem}: {episodes} episodes, T={T}, K={K}, terminated episodes={term}, "
              f"steps={int(packed['length'].sum())}, {os.path.getsize(path)/1024:.0f} KiB")


MAZE_MAP = os.path.join(rh.REFERENCE_ROOT, "tests", "assets", "board_maze.txt")
CTF_MAP = os.path.join(rh.REFERENCE_ROOT, "tests", "assets", "board.txt")


def make_maze64(seed=0, size=64, density=0.2):
    """SURVEY 8(d) config 4 map: obstacle border, random obstacles (density 0.2), one flag; MazeWorld codes."""
    rng = np.random.default_rng(seed)
    m = (rng.random((size, size)) < density).astype(np.int64) * 3
    m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = 3
    free = np.argwhere(m == 0)
    fx, fy = free[rng.integers(len(free))]
    m[fx, fy] = 2
    return m


def gen_maze():
    import tempfile
    plans = [("maze_board13", MAZE_MAP, 0.0, 24), ("maze_board13_penalty", MAZE_MAP, 0.5, 24)]
    tmp = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    np.savetxt(tmp.name, make_maze64().T, fmt="%d")   # load_text_map transposes (utils/map.py:37)
    plans += [("maze_gen64", tmp.name, 0.0, 6), ("maze_gen64_penalty", tmp.name, 0.5, 6)]
    for stem, path, pen, episodes in plans:
        eps = [rh.record_maze_episode(path, seed, np.random.default_rng(2000 + seed), pen) for seed in range(episodes)]
        for e in eps:
            assert e["obs"].dtype == np.float64   # the reference's dtype (maze.py:246); values are small integers
            e["obs"] = e["obs"].astype(np.uint8)
            e["init_obs"] = e["init_obs"].astype(np.uint8)
            e["start_index"] = np.array(e["start_index"], np.int32)
        out = rh.pack_episodes(eps, ["actions", "obs", "reward", "terminated", "truncated", "pos", "dir", "info"],
                               ["field_map", "init_obs", "start_index", "init_info"])
        out["field_map"] = out["field_map"][0].astype(np.uint8)
        out["meta_obstacle_penalty_ratio"] = np.array(pen)
        out["meta_ref_obs_dtype"] = np.array("float64")
        path_out = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path_out, **out)
        print(f"{stem}: {episodes} episodes, steps={int(out['length'].sum())}, terminated={int(out['terminated'].any(1).sum())}, "
              f"{os.path.getsize(path_out)/1024:.0f} KiB")
    os.unlink(tmp.name)


class MovingActions:
    """Blue actions that never `stay`: with a collision penalty staying is a self-collision that ends the agent (ctf.py:1231-1236),
    so uniformly random episodes are over before two agents ever meet.  Moving blue agents live long enough to battle."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def integers(self, lo, hi, size=None):
        return self.rng.integers(max(lo, 1), hi, size=size)


def gen_ctf(only_penalty_battles=False):
    plans = [("ctf_2v2", 2, 2, 0.0, 32), ("ctf_3v4", 3, 4, 0.0, 12), ("ctf_2v2_penalty", 2, 2, 0.5, 12), ("ctf_1v1", 1, 1, 0.0, 8)]
    if only_penalty_battles:
        plans = []
    # round 2: collision penalty AND battles in the same episodes (ctf.py:1316-1332 next to :1359-1420): 3v4, blue never stays,
    # keeping only the episodes of 400 in which at least one battle was fought
    plans.append(("ctf_3v4_penalty_battles", 3, 4, 0.5, 400))
    for stem, nb, nr, pen, episodes in plans:
        if stem == "ctf_3v4_penalty_battles":
            eps = [rh.record_ctf_mvn_episode(CTF_MAP, seed, MovingActions(3300 + seed), nb, nr, pen) for seed in range(episodes)]
            eps = [e for e in eps if int(np.sum(e["n_battles"])) > 0]
            episodes = len(eps)
        else:
            eps = [rh.record_ctf_mvn_episode(CTF_MAP, seed, np.random.default_rng(3000 + seed), nb, nr, pen) for seed in range(episodes)]
        for e in eps:
            assert e["obs"].dtype == np.int64     # the reference's dtype (ctf.py:1138)
            e["obs"] = e["obs"].astype(np.uint8)
            e["init_obs"] = e["init_obs"].astype(np.uint8)
        out = rh.pack_episodes(eps, ["actions", "red_actions", "order", "n_battles", "blue_win", "obs", "reward", "terminated",
                                     "truncated", "pos", "dir", "dead", "info", "stats_flags", "stats_defeated"],
                               ["field_map", "init_obs", "init_pos", "init_dir", "blue_place", "red_place", "init_info"])
        out["field_map"] = out["field_map"][0].astype(np.uint8)
        out["meta_num_blue"], out["meta_num_red"] = np.array(nb), np.array(nr)
        out["meta_obstacle_penalty_ratio"] = np.array(pen)
        out["meta_ref_obs_dtype"] = np.array("int64")
        path_out = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path_out, **out)
        print(f"{stem}: {episodes} episodes, steps={int(out['length'].sum())}, battles={int(out['n_battles'].sum())}, "
              f"terminated={int(out['terminated'].any(1).sum())}, {os.path.getsize(path_out)/1024:.0f} KiB")


def gen_ctf_carry():
    """Consecutive episodes of ONE env instance: the reference's reset() keeps Agent.terminated / collided, so agents
    defeated in an earlier episode start the next one defeated.  Sessions x episodes, episode index = session * K + k; short
    max_steps keeps the file small and makes several resets per session."""
    for stem, nb, nr, pen, sessions, K, max_steps in (("ctf_2v2_carry", 2, 2, 0.0, 10, 6, 30), ("ctf_3v4_penalty_carry", 3, 4, 0.5, 6, 5, 30)):
        eps = []
        for sidx in range(sessions):
            acts = MovingActions(3900 + sidx) if pen else np.random.default_rng(3900 + sidx)
            eps += rh.record_ctf_mvn_session(CTF_MAP, sidx, acts, K, nb, nr, pen, max_steps=max_steps)
        for e in eps:
            assert e["obs"].dtype == np.int64     # the reference's dtype (ctf.py:1138)
            e["obs"] = e["obs"].astype(np.uint8)
            e["init_obs"] = e["init_obs"].astype(np.uint8)
        out = rh.pack_episodes(eps, ["actions", "red_actions", "order", "n_battles", "blue_win", "obs", "reward", "terminated",
                                     "truncated", "pos", "dir", "dead", "collided", "info", "stats_flags", "stats_defeated"],
                               ["field_map", "init_obs", "init_pos", "init_dir", "init_dead", "init_collided", "blue_place", "red_place",
                                "init_info"])
        out["field_map"] = out["field_map"][0].astype(np.uint8)
        out["meta_num_blue"], out["meta_num_red"] = np.array(nb), np.array(nr)
        out["meta_obstacle_penalty_ratio"] = np.array(pen)
        out["meta_sessions"], out["meta_episodes_per_session"], out["meta_max_steps"] = np.array(sessions), np.array(K), np.array(max_steps)
        path_out = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path_out, **out)
        print(f"{stem}: {sessions} sessions x {K} episodes, steps={int(out['length'].sum())}, battles={int(out['n_battles'].sum())}, "
              f"episodes starting with a defeated agent={int(out['init_dead'].any(1).sum())}, with a collided one="
              f"{int(out['init_collided'].any(1).sum())}, {os.path.getsize(path_out)/1024:.0f} KiB")


def gen_ctf_render():
    """CtF episodes with the agents' sticky background colour after every step and rgb_array frames every 7th step (and at the
    end): the replay inputs of gen_ctf plus what `render()` shows.  Penalty 0.5 episodes add collided (grey) agents."""
    for stem, nb, nr, pen, episodes in (("render_ctf_2v2", 2, 2, 0.0, 5), ("render_ctf_3v4_penalty", 3, 4, 0.5, 4)):
        eps = [rh.record_ctf_mvn_episode(CTF_MAP, seed, np.random.default_rng(3500 + seed), nb, nr, pen, render_every=7) for seed in range(episodes)]
        for e in eps:
            e["obs"] = e["obs"].astype(np.uint8)
            e["init_obs"] = e["init_obs"].astype(np.uint8)
        out = rh.pack_episodes(eps, ["actions", "red_actions", "order", "n_battles", "blue_win", "obs", "pos", "dir", "dead", "bg"],
                               ["field_map", "init_obs", "init_pos", "init_dir", "init_bg", "blue_place", "red_place"])
        out["field_map"] = out["field_map"][0].astype(np.uint8)
        out["meta_num_blue"], out["meta_num_red"] = np.array(nb), np.array(nr)
        out["meta_obstacle_penalty_ratio"] = np.array(pen)
        out["frame_episode"] = np.concatenate([np.full(len(e["frame_step"]), i, np.int32) for i, e in enumerate(eps)])
        out["frame_step"] = np.concatenate([e["frame_step"] for e in eps])       # -1 = after reset
        for ts in (32, 8):
            out[f"frames_{ts}"] = np.concatenate([e[f"frames_{ts}"] for e in eps])
        path_out = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path_out, **out)
        print(f"{stem}: {episodes} episodes, steps={int(out['length'].sum())}, frames={len(out['frame_step'])}, dead agents drawn="
              f"{int(sum(out['dead'][e, s].sum() for e, s in zip(out['frame_episode'], out['frame_step']) if s >= 0))}, "
              f"bg != team colour on {int((out['bg'] != out['init_bg'][:, None]).sum())} agent-steps, {os.path.getsize(path_out)/1024:.0f} KiB")


def gen_ctf_flat():
    """observation_option="flattened" (ctf.py:1084-1104) - what the reference's own RL script feeds its policy
    (scripts/main_mvn_ctf_rl.py:15-21): the replay inputs of gen_ctf with the flattened int64 vectors as `obs`."""
    for stem, nb, nr, episodes in (("ctf_2v2_flat", 2, 2, 6), ("ctf_3v4_flat", 3, 4, 3)):
        eps = [rh.record_ctf_mvn_episode(CTF_MAP, seed, np.random.default_rng(3700 + seed), nb, nr, 0.0, observation_option="flattened")
               for seed in range(episodes)]
        for e in eps:
            assert e["obs"].dtype == np.int64 and e["obs"].ndim == 2 and e["init_obs"].dtype == np.int64
            assert e["obs"].max() < 256 and e["obs"].min() >= 0
            e["obs"] = e["obs"].astype(np.uint8)            # stored compactly; the reference's dtype is int64 (meta_ref_obs_dtype)
            e["init_obs"] = e["init_obs"].astype(np.uint8)
        out = rh.pack_episodes(eps, ["actions", "red_actions", "order", "n_battles", "blue_win", "obs", "pos", "dead"],
                               ["field_map", "init_obs", "init_pos", "blue_place", "red_place"])
        out["field_map"] = out["field_map"][0].astype(np.uint8)
        out["meta_num_blue"], out["meta_num_red"] = np.array(nb), np.array(nr)
        out["meta_obstacle_penalty_ratio"] = np.array(0.0)
        out["meta_ref_obs_dtype"] = np.array("int64")
        path_out = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path_out, **out)
        print(f"{stem}: {episodes} episodes, steps={int(out['length'].sum())}, flattened length {out['obs'].shape[-1]}, {os.path.getsize(path_out)/1024:.0f} KiB")
    # Ctf1v1Env (ctf.py:359-371): same layout, but the tail is the single `is_red_agent_defeated` flag
    eps = [rh.record_ctf_1v1_episode(CTF_MAP, seed, np.random.default_rng(3800 + seed), observation_option="flattened") for seed in range(8)]
    for e in eps:
        assert e["obs"].dtype == np.int64 and e["obs"].max() < 256
        e["obs"] = e["obs"].astype(np.uint8)
        e["init_obs"] = e["init_obs"].astype(np.uint8)
    out = rh.pack_episodes(eps, ["actions", "red_actions", "n_battles", "blue_win", "obs", "pos", "dead"],
                           ["field_map", "init_obs", "init_pos", "blue_place", "red_place"])
    out["field_map"] = out["field_map"][0].astype(np.uint8)
    path_out = os.path.join(OUT, "ctf1v1_flat.npz")
    np.savez_compressed(path_out, **out)
    print(f"ctf1v1_flat: {len(eps)} episodes, steps={int(out['length'].sum())}, flattened length {out['obs'].shape[-1]}, "
          f"red defeated on {int(out['dead'][..., 1].sum())} steps, {os.path.getsize(path_out)/1024:.0f} KiB")


def gen_ctf1v1():
    eps = [rh.record_ctf_1v1_episode(CTF_MAP, seed, np.random.default_rng(4000 + seed)) for seed in range(40)]
    for e in eps:
        assert e["obs"].dtype == np.int64
        e["obs"] = e["obs"].astype(np.uint8)
        e["init_obs"] = e["init_obs"].astype(np.uint8)
    out = rh.pack_episodes(eps, ["actions", "red_actions", "n_battles", "blue_win", "obs", "reward", "terminated", "truncated",
                                 "pos", "dir", "dead", "info", "stats_flags", "stats_defeated"],
                           ["field_map", "init_obs", "init_pos", "blue_place", "red_place", "init_info"])
    out["field_map"] = out["field_map"][0].astype(np.uint8)
    path_out = os.path.join(OUT, "ctf1v1.npz")
    np.savez_compressed(path_out, **out)
    print(f"ctf1v1: {len(eps)} episodes, steps={int(out['length'].sum())}, battles={int(out['n_battles'].sum())}, "
          f"terminated={int(out['terminated'].any(1).sum())}, {os.path.getsize(path_out)/1024:.0f} KiB")


def gen_partial():
    rh.import_reference()
    parts = [rh.record_partial_views("multigrid-collect-rooms-respawn-v0", 11, 120),
             rh.record_partial_views("multigrid-collect-respawn-clustered-v0", 12, 120),
             rh.record_partial_views("multigrid-collect-quadrants15-v0", 13, 60)]
    for stem, r in zip(("partial_rooms", "partial_clustered", "partial_quadrants15"), parts):
        path = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path, **r)
        print(f"{stem}: {len(r['V'])} states x {r['pos'].shape[1]} agents, V in {sorted(set(r['V'].tolist()))}, "
              f"see_through in {sorted(set(r['see_through'].tolist()))}, {os.path.getsize(path)/1024:.0f} KiB")


def gen_toroid():
    for stem, env_id, n in (("toroid_clustered", "multigrid-collect-respawn-clustered-v0", 40),
                            ("toroid_rooms", "multigrid-collect-rooms-respawn-v0", 40),
                            ("toroid_single", "multigrid-collect-single-v0", 20)):
        r = rh.record_toroid(env_id, 21, n)
        path = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path, **r)
        print(f"{stem}: {n} states, toroid {r['toroid'].shape} {r['toroid'].dtype}, {os.path.getsize(path)/1024:.0f} KiB")


def gen_render():
    for stem, env_id, n in (("render_clustered", "multigrid-collect-respawn-clustered-v0", 4),
                            ("render_rooms", "multigrid-collect-rooms-respawn-v0", 3),
                            ("render_quadrants15", "multigrid-collect-quadrants15-v0", 2)):
        r = rh.record_render(env_id, 41, n)
        path = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path, **r)
        print(f"{stem}: frames {r['frames_32'].shape} + {r['frames_8'].shape}, {os.path.getsize(path)/1024:.0f} KiB")


def gen_maze_render():
    r = rh.record_maze_render(MAZE_MAP, 51, 8)
    path = os.path.join(OUT, "render_maze13.npz")
    np.savez_compressed(path, **r)
    print(f"render_maze13: frames {r['frames_32'].shape} + {r['frames_8'].shape}, dirs {sorted(set(r['dir'].tolist()))}, {os.path.getsize(path)/1024:.0f} KiB")


def gen_generic_partial():
    for stem, size, A, n in (("partial6_9x9_a3", 9, 3, 120), ("partial6_12x12_a5", 12, 5, 60)):
        r = rh.record_generic_partial(size, A, 31, n)
        r["meta_size"], r["meta_num_agents"] = np.array(size), np.array(A)
        path = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path, **r)
        print(f"{stem}: {len(r['V'])} states x {A} agents, V in {sorted(set(r['V'].tolist()))}, "
              f"see_through in {sorted(set(r['see_through'].tolist()))}, {os.path.getsize(path)/1024:.0f} KiB")


def gen_generic():
    """Base-class MultiGridEnv.step + encode_dim-6 encode_for_agents on a DefaultWorld env assembled from reference classes."""
    for stem, size, A, max_steps, episodes in (("generic_9x9_a3", 9, 3, 40, 24), ("generic_12x12_a5", 12, 5, 60, 12),
                                                ("generic_7x7_a1", 7, 1, 30, 8)):
        eps = [rh.record_generic_episode(size, A, max_steps, seed, np.random.default_rng(5000 + seed)) for seed in range(episodes)]
        out = rh.pack_episodes(eps, ["actions", "order", "obs", "rewards", "terminated", "truncated", "pos", "dir"],
                               ["init_obs", "init_pos", "init_dir"])
        out["meta_size"], out["meta_num_agents"], out["meta_max_steps"] = np.array(size), np.array(A), np.array(max_steps)
        path_out = os.path.join(OUT, stem + ".npz")
        np.savez_compressed(path_out, **out)
        print(f"{stem}: {episodes} episodes, steps={int(out['length'].sum())}, terminated={int(out['terminated'].any(1).sum())}, "
              f"{os.path.getsize(path_out)/1024:.0f} KiB")


def gen_policies():
    """Decisions of the reference's scripted CtF opponents (policy/ctf/heuristic.py) and routes of its A* (policy/ctf/utils.py)
    on recorded inputs.  Every policy x team is driven for a few hundred decisions with ONE seeded Generator, partly closed
    loop (the agent makes the move it chose), so the actions pin the targets, the routes' tie-breaking and the order and kind of
    every random draw; `tail` is a draw made after the last decision (the stream position)."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    from replay import POLICY_NAMES, policy_maps, policy_observation
    rh.import_reference()
    from gym_multigrid.policy.ctf import heuristic as H
    from gym_multigrid.policy.ctf.utils import a_star
    maps = policy_maps()
    out = {}
    rng = np.random.default_rng(2024)
    for name, fm in maps.items():
        rows, cols = fm.shape
        n = 260 if name != "walls" else 420
        starts = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1)
        ends = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1)
        ends[:6] = starts[:6]                                         # start == end
        paths = [np.array(a_star(tuple(s), tuple(e), fm), np.int64).reshape(-1, 2) for s, e in zip(starts, ends)]
        out[f"astar_{name}_start"], out[f"astar_{name}_end"] = starts, ends
        out[f"astar_{name}_len"] = np.array([len(p) for p in paths])
        out[f"astar_{name}_cells"] = np.concatenate(paths)
        print(f"a_star {name}: {n} routes, {sum(len(p) == 0 for p in paths)} without a route, longest {max(len(p) for p in paths)}")
    moves = {0: (0, 0), 1: (0, -1), 2: (-1, 0), 3: (0, 1), 4: (1, 0)}
    seed = 9000
    for mname in ("board", "wide"):
        fm = maps[mname]
        rows, cols = fm.shape
        for pname in POLICY_NAMES:
            for ego in ("red", "blue"):
                seed += 1
                gen = np.random.Generator(np.random.PCG64(seed))
                kw = dict(field_map=fm, random_generator=gen, ego_agent=ego, randomness=0.75 if ego == "red" else 0.6)
                pol = getattr(H, pname)(**kw)
                K, nb, nr = 320, 3, 2
                cur = np.array([rng.integers(0, rows), rng.integers(0, cols)])
                rec = dict(curr=[], blue=[], red=[], action=[])
                blue = np.stack([rng.integers(0, rows, nb), rng.integers(0, cols, nb)], 1)
                red = np.stack([rng.integers(0, rows, nr), rng.integers(0, cols, nr)], 1)
                for k in range(K):
                    if k % 40 == 0:                                   # a jump: fresh positions (on the border half of the time)
                        border = getattr(pol, "border", [])
                        cur = np.array(border[rng.integers(0, len(border))]) if len(border) and rng.random() < 0.5 else \
                            np.array([rng.integers(0, rows), rng.integers(0, cols)])
                    step = np.array([moves[int(a)] for a in rng.integers(0, 5, nb + nr)])
                    blue = np.clip(blue + step[:nb], 0, [rows - 1, cols - 1])
                    red = np.clip(red + step[nb:], 0, [rows - 1, cols - 1])
                    (red if ego == "red" else blue)[0] = cur          # the controlled agent is the first of its team
                    a = int(pol.act(policy_observation(fm, blue, red), tuple(cur)))
                    for key, v in (("curr", cur), ("blue", blue), ("red", red), ("action", a)):
                        rec[key].append(np.array(v))
                    cur = np.clip(cur + moves[a], 0, [rows - 1, cols - 1])
                stem = f"{mname}_{pname}_{ego}"
                for key, v in rec.items():
                    out[f"{stem}_{key}"] = np.stack(v).astype(np.int16)
                out[f"{stem}_seed"] = np.array(seed)
                out[f"{stem}_tail"] = np.array(gen.integers(0, 2 ** 31))
                out[f"{stem}_randomness"] = np.array(kw["randomness"])
                if hasattr(pol, "border"):
                    out[f"{stem}_border"] = np.array(pol.border, np.int64).reshape(-1, 2)
                print(f"{stem}: {K} decisions, action histogram {np.bincount(out[f'{stem}_action'], minlength=5).tolist()}")
    path_out = os.path.join(OUT, "ctf_policies.npz")
    np.savez_compressed(path_out, **out)
    print(f"ctf_policies: {os.path.getsize(path_out)/1024:.0f} KiB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["collect", "maze", "ctf", "ctf_carry", "ctf1v1", "partial", "toroid", "generic", "generic_partial", "render", "ctf_flat", "policies"]
    if "collect" in which:
        gen_collect()
    elif "collect_variants" in which:      # only the COLLECT_VARIANTS fixtures (added in round 2; the others are unchanged)
        gen_collect(only_variants=True)
    if "maze" in which:
        gen_maze()
    if "ctf" in which:
        gen_ctf()
    elif "ctf_penalty_battles" in which:
        gen_ctf(only_penalty_battles=True)
    if "ctf_carry" in which:               # round 2, closing pass: episodes of one env instance (flags surviving reset)
        gen_ctf_carry()
    if "ctf1v1" in which:
        gen_ctf1v1()
    if "partial" in which:
        gen_partial()
    if "toroid" in which:
        gen_toroid()
    if "ctf_flat" in which:
        gen_ctf_flat()
    if "render" in which:
        gen_render()
        gen_maze_render()
        gen_ctf_render()
    if "generic" in which:
        gen_generic()
    if "generic_partial" in which:
        gen_generic_partial()
    if "policies" in which:
        gen_policies()
