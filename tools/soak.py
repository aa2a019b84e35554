#!/usr/bin/env python
"""Randomised differential soak: CUDA (through the C ABI) against the CPU oracle on configurations drawn at random - sizes,
layouts, team sizes, maps, penalties, tile sizes, env counts the fixed tests never use.  Runs on the GPU box only.

    python tools/soak.py [--seconds 120] [--seed 0] > profiles/r01_soak.json

Every comparison is bit-exact (np.array_equal); the first mismatch aborts with the offending configuration.  The summary
(JSON, one line) counts configurations, env-steps and compared bytes per family.  The oracle is the checker here, as in tests/.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import gym_multigrid_b200 as mg  # noqa: E402
import oracle as oc  # noqa: E402
from gym_multigrid_b200.vector_env import CollectVecEnv  # noqa: E402

DEV = "cuda:0"


def _np(t):
    return t.cpu().numpy()


def same(a, b, what, cfg):
    if not np.array_equal(a, b):
        raise SystemExit(f"MISMATCH {what}: {cfg}")
    return a.nbytes


def soak_collect(rng, stats):
    layout = str(rng.choice(["even_dist", "quadrants", "quadrants_respawn", "rooms"]))
    A = int(rng.integers(1, 9)) if layout == "even_dist" else int(rng.integers(2, 4))
    nb = int(rng.integers(1, 5))
    size = int(rng.integers(7, 25)) if layout != "rooms" else int(rng.choice([11, 13, 15, 17]))
    if layout in ("quadrants", "quadrants_respawn"):
        size += size % 2
    colours = rng.permutation(10)
    kw = dict(size=size, agents_index=[int(c) for c in colours[:A]], balls_index=[int(c) for c in rng.permutation(10)[:nb]],
              balls_reward=[float(rng.choice([1, 2, -1, 0.5])) for _ in range(nb)], num_balls=int(rng.integers(nb, max(nb + 1, size))),
              respawn=bool(rng.integers(0, 2)) or layout == "quadrants_respawn", layout=layout)
    tl, n, seed, base = int(rng.integers(5, 40)), int(rng.integers(1, 400)), int(rng.integers(0, 1 << 30)), int(rng.integers(0, 1000))
    cfg = dict(family="collect", n=n, tl=tl, seed=seed, **kw)
    try:
        env = CollectVecEnv(n, max_episode_steps=tl, seed=seed, env_id_base=base, **kw)
    except (RuntimeError, ValueError):
        stats["collect_rejected"] += 1        # e.g. more objects than cells: both sides refuse such a config
        return
    env.enable_final_observation()
    o = oc.CollectOracle(oc.make_collect_cfg(time_limit=tl, **kw), n)
    r = oc.PhiloxRng(seed=seed, env_id_base=base)
    b = same(_np(env.reset()[0]), o.reset(r), "collect reset", cfg)
    steps = int(rng.integers(tl, 3 * tl))
    for t in range(steps):
        act = rng.integers(-1, 5, size=(n, A)).astype(np.int8)            # includes out-of-range actions (no-ops)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=DEV))
        oobs, orew, oterm, otrunc, ofin = o.step(act, r, autoreset=True, want_final_obs=True)
        b += same(_np(obs), oobs, f"collect obs step {t}", cfg) + same(_np(rew), orew, "collect rewards", cfg)
        b += same(_np(term), oterm, "collect terminated", cfg) + same(_np(trunc), otrunc, "collect truncated", cfg)
        d = oterm | otrunc
        b += same(_np(info["final_observation"])[d], ofin[d], "collect final obs", cfg)
    # ... continued as ONE launch (mg_rollout, state held in shared memory over T steps; drawn T includes 1)
    T = int(rng.integers(1, 2 * tl))
    racts = rng.integers(-1, 5, size=(T, n, A)).astype(np.int8)
    robs, rrew, rterm, rtrunc, rinfo = env.rollout(torch.as_tensor(racts, device=DEV), final_observation=True)
    robs, rrew, rterm, rtrunc, rfin = _np(robs), _np(rrew), _np(rterm), _np(rtrunc), _np(rinfo["final_observation"])
    for t in range(T):
        oobs, orew, oterm, otrunc, ofin = o.step(racts[t], r, autoreset=True, want_final_obs=True)
        b += same(robs[t], oobs, f"rollout obs step {t} of {T}", cfg) + same(rrew[t], orew, "rollout rewards", cfg)
        b += same(rterm[t], oterm, "rollout terminated", cfg) + same(rtrunc[t], otrunc, "rollout truncated", cfg)
        d = oterm | otrunc
        b += same(rfin[t][d], ofin[d], "rollout final obs", cfg)
    # ... and through host buffers with one of the result transports (expanded obs / packed plane / per-env delta records)
    tr = str(rng.choice(["full", "packed", "delta"]))
    env.set_host_transport(tr, int(rng.integers(1, 5)))
    for t in range(int(rng.integers(2, tl + 4))):
        act = rng.integers(-1, 5, size=(n, A)).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(act)
        oobs, orew, oterm, otrunc = o.step(act, r, autoreset=True)
        b += same(obs, oobs, f"host step {t} obs ({tr})", cfg) + same(rew, orew, f"host rewards ({tr})", cfg)
        b += same(term, oterm, f"host terminated ({tr})", cfg) + same(trunc, otrunc, f"host truncated ({tr})", cfg)
    steps += T
    stats["collect_rollout_steps"] += n * T; stats[f"collect_host_{tr}_configs"] += 1
    b += same(_np(env.grid), o.grid, "collect grid", cfg) + same(_np(env.pickups).reshape(n, -1), o.info, "collect counters", cfg)
    dirs = rng.integers(0, 4, size=(n, A)).astype(np.uint8)
    V, st = int(rng.choice([3, 4, 5, 6, 7, 9])), bool(rng.integers(0, 2))
    b += same(_np(env.gen_obs(V, st, dirs=dirs)), oc.partial_view3(o.grid, o.agent_pos, size, size, V, st, dirs=dirs), "collect views", cfg)
    b += same(_np(env.toroid_obs()), oc.toroid(o.grid, o.agent_pos, size, nb), "toroid", cfg)
    ts = int(rng.choice([32, 16, 8, 5, 3]))
    ids = rng.integers(0, n, size=min(n, 6))
    b += same(_np(env.render(env_ids=ids, tile_size=ts)), oc.render_grid(oobs[ids], ts), "collect render", cfg)
    assert env.status() == 0
    env.close()
    stats["collect_configs"] += 1; stats["collect_env_steps"] += n * steps; stats["bytes_compared"] += b


def random_ctf_map(rng, S):
    """Left half blue territory, right half red, random obstacles, one flag each (CtfWorld codes, field_map[x, y])."""
    m = np.zeros((S, S), np.uint8)
    m[:, S // 2:] = 1
    m[rng.random((S, S)) < 0.08] = 6
    bx, by = int(rng.integers(0, S)), int(rng.integers(0, S // 2))
    rx, ry = int(rng.integers(0, S)), int(rng.integers(S // 2, S))
    m[bx, by], m[rx, ry] = 4, 5
    return m


def soak_ctf(rng, stats):
    S = int(rng.integers(5, 21))
    fm = random_ctf_map(rng, S)
    v1 = bool(rng.integers(0, 5) == 0)
    nb, nr = (1, 1) if v1 else (int(rng.integers(1, 9)), int(rng.integers(1, 9)))
    if (fm == 0).sum() + 1 < nb or (fm == 1).sum() + 1 < nr:
        return
    pen = 0.0 if v1 else float(rng.choice([0.0, 0.0, 0.5]))
    kw = dict(battle_range=float(rng.choice([1.0, 1.5, 2.0, 3.0])), randomness=float(rng.choice([0.75, 0.5, 0.9])),
              max_steps=int(rng.integers(5, 50)))
    n, seed, base = int(rng.integers(1, 700)), int(rng.integers(0, 1 << 30)), int(rng.integers(0, 1000))
    ref = bool(rng.integers(0, 4) == 0)
    cfg = dict(family="ctf1v1" if v1 else "ctf", S=S, nb=nb, nr=nr, pen=pen, n=n, seed=seed, ref_dtypes=ref, **kw)
    if v1:
        env = mg.make_ctf1v1_vec(n, fm, seed=seed, env_id_base=base, reference_dtypes=ref, **kw)
        o = oc.CtfOracle(fm, n, 1, 1, variant_1v1=True, **kw)
    else:
        env = mg.make_ctf_vec(n, fm, num_blue_agents=nb, num_red_agents=nr, obstacle_penalty_ratio=pen, seed=seed, env_id_base=base, reference_dtypes=ref, **kw)
        o = oc.CtfOracle(fm, n, nb, nr, obstacle_penalty_ratio=pen, **kw)
    mk = lambda **a: oc.map_rng(mode=1, seed=seed, env_id_base=base, **a)  # noqa: E731
    b = same(_np(env.reset()[0]), o.reset(mk()), "ctf reset", cfg)
    steps = int(rng.integers(10, 80))
    ext = bool(rng.integers(0, 3) == 0)
    red = env.set_red_actions(torch.zeros((n, nr), dtype=torch.int8, device=DEV)) if ext else None
    tables = None
    if not ext and rng.integers(0, 3) == 0:      # the scripted opponents decided on the device (ctf_policy_kernel) vs the oracle's rule
        from gym_multigrid_b200.policy.ctf import heuristic as H
        fmf = fm.astype(np.float64)
        classes = [H.RwPolicy, H.FightPolicy, H.CapturePolicy, H.PatrolPolicy, H.PatrolFightPolicy]
        picks = [classes[int(rng.integers(0, 5))] for _ in range(nr)]
        pols = [c() if c is H.RwPolicy else c(fmf, randomness=float(rng.choice([0.75, 0.3, 1.0]))) for c in picks]
        try:
            env.set_enemy_policies(pols, device=True, fused=bool(rng.integers(0, 4) != 0))   # inside the step kernel / explicit launch before it
        except ValueError:                        # e.g. a map whose territories do not touch: no border to patrol
            pass
        if env._device_policies:
            tables = env._policy_tables
            cfg["device_policies"] = [c.__name__ for c in picks]
            cfg["fused_policies"] = env._fused_policies
            stats["ctf_device_policy_configs"] += 1
            stats["ctf_fused_policy_configs"] = stats.get("ctf_fused_policy_configs", 0) + int(env._fused_policies)
    for t in range(steps):
        act = rng.integers(0, 5, size=(n, nb)).astype(np.int8)
        ra = None
        if ext:
            ra = rng.integers(0, 5, size=(n, nr)).astype(np.int8)
            red.copy_(torch.as_tensor(ra))
        elif tables is not None:
            ra = o.policy_actions(tables, seed, _np(env.episode_count), env_id_base=base)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=DEV))
        if tables is not None:
            b += same(_np(env._red_buf), ra, f"ctf device policy actions step {t}", cfg)
        oobs, orew, oterm, otrunc = o.step(act, mk(red_actions=ra) if ra is not None else mk(), autoreset=True)
        b += same(_np(obs), oobs, f"ctf obs step {t}", cfg) + same(_np(rew), orew, "ctf reward", cfg)
        b += same(_np(term), oterm, "ctf terminated", cfg) + same(_np(trunc), otrunc, "ctf truncated", cfg)
    b += same(_np(env.agent_pos), o.pos, "ctf pos", cfg) + same(_np(env.agent_flags), o.flags, "ctf flags", cfg)
    b += same(_np(env.flattened_obs()), o.flattened(), "ctf flattened", cfg)
    b += same(np.stack([_np(v) for v in env.get_info().values()], 1), o.info(), "ctf info", cfg)
    ts = int(rng.choice([32, 8, 6, 3]))
    ids = rng.integers(0, n, size=min(n, 5))
    b += same(_np(env.render(env_ids=ids, tile_size=ts)), oc.render_ctf(fm, o.pos[ids], o.dir[ids], o.flags[ids], nb, ts, variant_1v1=v1), "ctf render", cfg)
    assert env.status() == 0
    env.close()
    stats["ctf_configs"] += 1; stats["ctf_env_steps"] += n * steps; stats["bytes_compared"] += b


def soak_maze(rng, stats):
    S = int(rng.integers(4, 70))
    fm = (rng.random((S, S)) < float(rng.choice([0.0, 0.1, 0.25]))).astype(np.uint8) * 3
    fm[int(rng.integers(0, S)), int(rng.integers(0, S))] = 2
    if (fm == 0).sum() == 0:
        return
    pen = float(rng.choice([0.0, 0.5]))
    n, seed, base, ms = int(rng.integers(1, 500)), int(rng.integers(0, 1 << 30)), int(rng.integers(0, 1000)), int(rng.integers(5, 60))
    ref = bool(rng.integers(0, 4) == 0) and S <= 32
    cfg = dict(family="maze", S=S, pen=pen, n=n, seed=seed, max_steps=ms, ref_dtypes=ref)
    env = mg.make_maze_vec(n, fm, obstacle_penalty_ratio=pen, max_steps=ms, seed=seed, env_id_base=base, reference_dtypes=ref)
    o = oc.MazeOracle(fm, n, obstacle_penalty_ratio=pen, max_steps=ms)
    mk = lambda: oc.map_rng(mode=1, seed=seed, env_id_base=base)  # noqa: E731
    b = same(_np(env.reset()[0]), o.reset(mk()), "maze reset", cfg)
    steps = int(rng.integers(10, 80))
    for t in range(steps):
        act = rng.integers(0, 5, size=n).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=DEV))
        oobs, orew, oterm, otrunc = o.step(act, mk(), autoreset=True)
        b += same(_np(obs), oobs, f"maze obs step {t}", cfg) + same(_np(rew), orew, "maze reward", cfg)
        b += same(_np(term), oterm, "maze terminated", cfg) + same(_np(trunc), otrunc, "maze truncated", cfg)
    b += same(np.stack([_np(v) for v in env.get_info().values()], 1), o.info(), "maze info", cfg)
    ts = int(rng.choice([8, 4, 3]))
    ids = rng.integers(0, n, size=min(n, 3))
    a = _np(env._agents)[:, 0]
    b += same(_np(env.render(env_ids=ids, tile_size=ts)), oc.render_maze(fm, a[ids, :2].astype(np.int16), a[ids, 2].astype(np.int8), ts), "maze render", cfg)
    assert env.status() == 0
    env.close()
    stats["maze_configs"] += 1; stats["maze_env_steps"] += n * steps; stats["bytes_compared"] += b


def soak_wildfire(rng, stats):
    """The Wildfire extension against the in-repo oracle of its specification (no reference code exists: parity unpinned)."""
    W = int(rng.choice([4, 8, 12, 16, 20, 24, 32, 40, 64, 96]))
    H = W if rng.integers(0, 3) else int(rng.choice([4, 6, 8, 10, 12, 16, 20, 32]))
    if (W * H) % 16:
        return
    A = int(rng.integers(1, min(32, W * H // 2) + 1))
    fires = int(rng.integers(1, min(8, W * H - A) + 1))
    n, seed, base, ms = int(rng.integers(1, 200)), int(rng.integers(0, 1 << 30)), int(rng.integers(0, 1000)), int(rng.integers(5, 50))
    kw = dict(num_agents=A, num_fires=fires, alpha=float(rng.choice([0.05, 0.2, 0.5])), beta=float(rng.choice([0.02, 0.08, 0.3])),
              max_steps=ms, seed=seed, env_id_base=base)
    kw.update(dict(size=W) if W == H else dict(width=W, height=H))
    cfg = dict(family="wildfire", W=W, H=H, A=A, fires=fires, n=n, **{k: kw[k] for k in ("alpha", "beta", "max_steps", "seed")})
    env = mg.make_wildfire_vec(n, **kw)
    env.enable_final_observation()
    o = oc.WildfireOracle(n, **kw)
    b = same(_np(env.reset()[0]), o.reset(), "wildfire reset", cfg)
    steps = int(rng.integers(10, 70))
    for t in range(steps):
        act = rng.integers(0, 5, size=(n, A)).astype(np.int8)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=DEV))
        oo, orew, oterm, otrunc, ofin = o.step(act, autoreset=True, want_final_obs=True)
        b += same(_np(obs), oo, f"wildfire obs step {t}", cfg) + same(_np(rew), orew, "wildfire rewards", cfg)
        b += same(_np(term), oterm, "wildfire terminated", cfg) + same(_np(trunc), otrunc, "wildfire truncated", cfg)
        d = oterm | otrunc
        b += same(_np(info["final_observation"])[d], ofin[d], "wildfire final observation", cfg)
    b += same(_np(env.terrain), o.terrain, "wildfire terrain", cfg) + same(_np(env.agents), o.agents, "wildfire agents", cfg)
    env.close()
    stats["wildfire_configs"] += 1; stats["wildfire_env_steps"] += n * steps; stats["bytes_compared"] += b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    import collections
    stats = collections.defaultdict(int)
    for k in ("collect_configs", "collect_env_steps", "collect_rejected", "ctf_configs", "ctf_env_steps", "maze_configs", "maze_env_steps",
              "bytes_compared", "ctf_device_policy_configs", "collect_rollout_steps"):
        stats[k] = 0
    t0 = time.time()
    fams = [soak_collect, soak_ctf, soak_maze, soak_wildfire]
    i = 0
    while time.time() - t0 < args.seconds:
        fams[i % len(fams)](rng, stats)
        i += 1
    stats.update(seconds=round(time.time() - t0, 1), seed=args.seed, mismatches=0,
                 gpu=torch.cuda.get_device_name(0), what="CUDA (C ABI) vs CPU oracle, bit-exact, random configurations")
    print(json.dumps(dict(stats)))


if __name__ == "__main__":
    main()
