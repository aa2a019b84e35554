#!/usr/bin/env python
"""Runs a few launches of ONE kernel family so that `ncu -k regex:<kernel> -s 3 -c 1 python tools/profile_families.py <family>`
can capture it (development aid; maps come from the committed golden fixtures)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg  # noqa: E402


def golden(stem, key):
    with np.load(os.path.join(ROOT, "tests", "golden", stem + ".npz")) as z:
        return z[key]


def main():
    fam = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = "cuda:0"
    if fam == "ctf":
        n = n or (1 << 20)
        e = mg.make_ctf_vec(n, golden("ctf_2v2", "field_map"))
        a = torch.randint(0, 5, (n, 2), device=dev, dtype=torch.int8)
        f = lambda: e.step(a)  # noqa: E731
    elif fam == "ctf8":             # 8v8: the general step body (run-time team sizes)
        n = n or (1 << 18)
        e = mg.make_ctf_vec(n, golden("ctf_2v2", "field_map"), num_blue_agents=8, num_red_agents=8)
        acts = [torch.randint(0, 5, (n, 8), device=dev, dtype=torch.int8) for _ in range(8)]
        it = iter(range(10 ** 9))
        f = lambda: e.step(acts[next(it) % 8])  # noqa: E731
    elif fam == "ctf_policy":      # scripted opponents decided on the device: ctf_policy_kernel ahead of every step
        from gym_multigrid_b200.policy.ctf.heuristic import FightPolicy, PatrolFightPolicy
        n = n or (1 << 20)
        fm = golden("ctf_2v2", "field_map")
        e = mg.make_ctf_vec(n, fm)
        e.set_enemy_policies([FightPolicy(fm.astype(np.float64)), PatrolFightPolicy(fm.astype(np.float64))], device=True)
        a = torch.randint(0, 5, (n, 2), device=dev, dtype=torch.int8)
        f = lambda: e.step(a)  # noqa: E731
    elif fam == "maze":
        n = n or 131072
        e = mg.make_maze_vec(n, golden("maze_gen64", "field_map"))
        a = torch.randint(0, 5, (n,), device=dev, dtype=torch.int8)
        f = lambda: e.step(a)  # noqa: E731
    elif fam == "maze_partial":     # BASELINE config 4: fused step + V = 7 partial views
        n = n or (1 << 20)
        e = mg.make_maze_vec(n, golden("maze_gen64", "field_map"))
        e.set_partial_obs(7)
        a = torch.randint(0, 5, (n,), device=dev, dtype=torch.int8)
        f = lambda: e.step(a)  # noqa: E731
    elif fam == "view_maze":
        n = n or 131072
        e = mg.make_maze_vec(n, golden("maze_gen64", "field_map"))
        out = torch.empty((n, 1, 7, 7, 3), dtype=torch.uint8, device=dev)
        f = lambda: e.gen_obs(7, False, out=out)  # noqa: E731
    elif fam in ("view_collect", "toroid"):
        n = n or 262144
        e = mg.make_vec("multigrid-collect-respawn-clustered-v0", n)
        out = torch.empty((n, 2, 7, 7, 3), dtype=torch.uint8, device=dev)
        tout = torch.empty((n, 2, 10, 10, 5), dtype=torch.float32, device=dev)
        f = (lambda: e.gen_obs(7, False, out=out)) if fam == "view_collect" else (lambda: e.toroid_obs(out=tout))  # noqa: E731
    elif fam == "wildfire":
        n = n or 16384
        e = mg.make_wildfire_vec(n, size=64, num_agents=16)
        acts = [torch.randint(0, 5, (n, 16), device=dev, dtype=torch.int8) for _ in range(12)]
        it = iter(range(10 ** 9))
        f = lambda: e.step(acts[next(it) % 12])  # noqa: E731
    elif fam == "render":
        n = n or 1024
        e = mg.make_vec("multigrid-collect-respawn-clustered-v0", n)
        out = torch.empty((n, 320, 320, 3), dtype=torch.uint8, device=dev)
        f = lambda: e.render(tile_size=32, out=out)  # noqa: E731
    elif fam == "generic":
        n = n or 65536
        g = {k: golden("generic_12x12_a5", k) for k in ("init_obs", "init_pos")}
        e = mg.make_generic_vec(n, 12, num_agents=5, max_steps=60)
        idx = np.arange(n) % g["init_obs"].shape[0]
        e.set_layout(g["init_obs"][idx, 0], g["init_pos"][idx])
        a = torch.randint(0, 4, (n, 5), device=dev, dtype=torch.int8)
        f = lambda: e.step(a)  # noqa: E731
    else:
        raise SystemExit(f"unknown family {fam}")
    e.reset()
    torch.cuda.synchronize()
    for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 12):
        f()
    torch.cuda.synchronize()
    e.close()


if __name__ == "__main__":
    main()
