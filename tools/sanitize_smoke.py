#!/usr/bin/env python
"""Small run of every kernel family for `compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_smoke.py`
(development aid: ragged sizes, autoreset, every observation mode; maps from the committed golden fixtures)."""
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
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    for env_id, n in (("multigrid-collect-respawn-clustered-v0", 333), ("multigrid-collect-rooms-respawn-v0", 130)):
        e = mg.make_vec(env_id, n, seed=1)
        e.enable_final_observation()
        e.reset()
        for _ in range(60):
            e.step(torch.randint(0, 4, (n, 2), generator=g, device=dev, dtype=torch.int8))
        for V in (3, 5, 7, 4, 9):
            e.gen_obs(V, False); e.gen_obs(V, True, dirs=torch.randint(0, 4, (n, 2), generator=g, device=dev, dtype=torch.uint8))
        e.toroid_obs(); e.encode()
        for ts in (32, 8, 5):
            e.render(tile_size=ts); e.render(env_ids=[n - 1, 0, 7], tile_size=ts)
        e.step_async(np.zeros((n, 2), np.int8)); e.step_wait()
        e.step(np.zeros((n, 2), np.int8))
        assert e.status() == 0
        e.close()
    fm = golden("ctf_2v2", "field_map")
    for nb, nr, ref in ((2, 2, False), (3, 4, False), (1, 1, True), (8, 8, False)):
        n = 257
        e = mg.make_ctf_vec(n, fm, num_blue_agents=nb, num_red_agents=nr, max_steps=20, reference_dtypes=ref)
        e.enable_final_observation()
        e.reset()
        for _ in range(45):
            e.step(torch.randint(0, 5, (n, nb), generator=g, device=dev, dtype=torch.int8))
        e.get_info()
        e.render(tile_size=8); e.render(env_ids=[3, n - 1], tile_size=3)
        red = e.set_red_actions(torch.randint(0, 5, (n, nr), generator=g, device=dev, dtype=torch.int8))
        e.step(torch.randint(0, 5, (n, nb), generator=g, device=dev, dtype=torch.int8))
        e.set_red_actions(None)
        assert e.status() == 0
        e.close()
    for stem in ("maze_board13", "maze_gen64"):
        fm = golden(stem, "field_map")
        for ref in (False, True):
            n = 200
            e = mg.make_maze_vec(n, fm, max_steps=15, reference_dtypes=ref)
            e.reset()
            for _ in range(35):
                e.step(torch.randint(0, 5, (n,), generator=g, device=dev, dtype=torch.int8))
            e.gen_obs(7); e.gen_obs(6); e.get_info()
            e.render(env_ids=[0, n - 1], tile_size=8); e.render(env_ids=[5], tile_size=7)
            if not ref:
                e.set_partial_obs(5)
                e.reset()
                for _ in range(20):
                    e.step(torch.randint(0, 5, (n,), generator=g, device=dev, dtype=torch.int8))
            assert e.status() == 0
            e.close()
    for kw in (dict(size=16, num_agents=5), dict(size=64, num_agents=16), dict(width=8, height=10, num_agents=6), dict(size=8, num_agents=32)):
        n = 70
        e = mg.make_wildfire_vec(n, num_fires=3, max_steps=12, **kw)
        e.enable_final_observation()
        e.reset()
        A = kw["num_agents"]
        for _ in range(30):
            e.step(torch.randint(0, 5, (n, A), generator=g, device=dev, dtype=torch.int8))
        e.close()
    gi = {k: golden("generic_12x12_a5", k) for k in ("init_obs", "init_pos")}
    n = 101
    e = mg.make_generic_vec(n, 12, num_agents=5, max_steps=9)
    idx = np.arange(n) % gi["init_obs"].shape[0]
    e.set_layout(gi["init_obs"][idx, 0], gi["init_pos"][idx])
    e.enable_final_observation()
    e.reset()
    for _ in range(25):
        e.step(torch.randint(0, 4, (n, 5), generator=g, device=dev, dtype=torch.int8))
    e.gen_obs(7); e.gen_obs(4, True)
    e.close()
    torch.cuda.synchronize()
    print("sanitize_smoke ok")


if __name__ == "__main__":
    main()
