#!/usr/bin/env python
"""The reference's scripted CtF opponents (tests/test_ctf.py:97-215) against this package: only the imports change.

    python examples/ctf_heuristic_opponents.py [--map path/to/board.txt] [--policy fight|capture|patrol|patrol_fight] [--num-envs 16]

* single env, reference style: `CtFMvNEnv(enemy_policies=[FightPolicy(), RwPolicy()])` - the policies decide on the host from the
  positional observation, exactly where the reference calls them; the step, the battles and the observation are CUDA kernels.
* a batch: `CtfVecEnv.set_enemy_policies(policy)` - the same objects for every env (a device sync per step; for very large
  batches pass `device=True`: the decisions move into a kernel, see below).
The policies reproduce the reference's decisions and random draws (tests/test_policies.py), A* tie-breaking included.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_multigrid_b200 as mg  # noqa: E402
from gym_multigrid_b200.map_env import load_text_map  # noqa: E402
from gym_multigrid_b200.policy.ctf.heuristic import CapturePolicy, FightPolicy, PatrolFightPolicy, PatrolPolicy, RwPolicy  # noqa: E402

POLICIES = {"fight": FightPolicy, "capture": CapturePolicy, "patrol": PatrolPolicy, "patrol_fight": PatrolFightPolicy}


def default_map(tmp="/tmp/board_b200.txt"):
    """tests/assets/board.txt of the reference, rebuilt from the golden fixture (`load_text_map` transposes, utils/map.py:37)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with np.load(os.path.join(root, "tests", "golden", "ctf_2v2.npz")) as z:
        np.savetxt(tmp, z["field_map"].T, fmt="%d")
    return tmp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default=None)
    ap.add_argument("--policy", default="fight", choices=sorted(POLICIES))
    ap.add_argument("--num-envs", type=int, default=16)
    args = ap.parse_args()
    map_path = args.map or default_map()
    field_map = load_text_map(map_path)

    # --- one env, the reference's test loop (tests/test_ctf.py:97-125)
    env = mg.CtFMvNEnv(num_blue_agents=2, num_red_agents=2, map_path=map_path, observation_option="flattened",
                       enemy_policies=[POLICIES[args.policy](field_map), RwPolicy()])
    obs, _ = env.reset()
    frames, total = [env.render()], 0.0
    while True:
        obs, reward, terminated, truncated, info = env.step(env.action_space.sample())
        frames.append(env.render())
        total += reward
        if terminated or truncated:
            break
    print(f"single env vs [{args.policy}, rw]: {len(frames) - 1} steps, return {total:+.2f}, frames {frames[0].shape} uint8")
    env.close()

    # --- a batch with the same opponents
    n = args.num_envs
    vec = mg.make_ctf_vec(n, map_path, num_blue_agents=2, num_red_agents=2, max_steps=100, seed=0)
    vec.set_enemy_policies(POLICIES[args.policy](field_map), random_generator=np.random.default_rng(0))
    vec.reset()
    ret, episodes = torch.zeros(n, dtype=torch.float64, device=vec.device), 0
    for _ in range(100):
        _, rew, term, trunc, _ = vec.step(torch.randint(0, 5, (n, 2), device=vec.device, dtype=torch.int8))
        ret += rew
        episodes += int((term | trunc).sum())
    print(f"{n} envs vs {args.policy} x 2: 100 steps, {episodes} episodes finished, mean return per env {float(ret.mean()):+.2f}")
    vec.close()

    # --- a large batch: the same opponents decided by a kernel (first moves of the reference's A* routes from a host-built table)
    n = 1 << 16
    big = mg.make_ctf_vec(n, map_path, num_blue_agents=2, num_red_agents=2, max_steps=100, seed=0)
    big.set_enemy_policies([POLICIES[args.policy](field_map), POLICIES[args.policy](field_map)], device=True)
    big.reset()
    ret = torch.zeros(n, dtype=torch.float64, device=big.device)
    for _ in range(100):
        _, rew, _, _, _ = big.step(torch.randint(0, 5, (n, 2), device=big.device, dtype=torch.int8))
        ret += rew
    print(f"{n} envs vs {args.policy} x 2 decided on the device: 100 steps, mean return per env {float(ret.mean()):+.2f}")
    big.close()


if __name__ == "__main__":
    main()
