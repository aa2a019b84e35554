#!/usr/bin/env python
"""Capture-the-Flag the way the reference's RL script uses it (scripts/main_mvn_ctf_rl.py): `observation_option="flattened"`
observations for the policy, rgb_array frames of a few envs for a GIF - batched, everything on one B200.

    python examples/ctf_policy_and_frames.py [--map path/to/board.txt] [--num-envs 4096] [--steps 100] [--out frames.npy]

* `CtfVecEnv(observation_option="flattened")`: `reset` / `step` return the vector of ctf.py:1084-1104 for every env (uint8; int64 with
  `reference_dtypes=True`).
* `set_red_actions(buffer)`: the red team follows whatever you write into `buffer` before each step (a learned opponent,
  self-play, a scripted policy) instead of the built-in random walk.
* `render(env_ids=[...])`: `MultiGridEnv.render()` frames of the selected envs, bit-identical to the reference's.
* `step_async` / `step_wait` (Collect shown in bench.py): the host-array path with two batches in flight.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_multigrid_b200 as mg  # noqa: E402


def default_map():
    """The reference's tests/assets/board.txt as recorded in the golden fixtures (field_map[x, y], CtfWorld codes)."""
    with np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ctf_2v2.npz")) as z:
        return z["field_map"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default=None)
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    env = mg.make_ctf_vec(args.num_envs, args.map or default_map(), num_blue_agents=2, num_red_agents=2, observation_option="flattened", seed=0)
    obs, _ = env.reset(seed=0)
    n, nb, nr = env.num_envs, env.num_blue, env.num_red
    red = env.set_red_actions(torch.zeros((n, nr), dtype=torch.int8, device=dev))
    blue_flag = torch.tensor(env.blue_flag, device=dev)
    frames, ret = [], torch.zeros(n, dtype=torch.float64, device=dev)
    for t in range(args.steps):
        # scripted opponent computed on the device from the state: step towards the blue flag, x first (CtfActions: 2 down = x-1, 4 up = x+1,
        # 1 left = y-1, 3 right = y+1)
        d = blue_flag - env.agent_pos[:, nb:].to(torch.int64)
        red.copy_(torch.where(d[..., 0] != 0, torch.where(d[..., 0] > 0, 4, 2), torch.where(d[..., 1] > 0, 3, torch.where(d[..., 1] < 0, 1, 0))).to(torch.int8))
        blue = torch.randint(0, 5, (n, nb), device=dev, dtype=torch.int8)          # your policy goes here: obs is [n, 216] on the GPU
        obs, rew, term, trunc, _ = env.step(blue)
        ret += rew
        frames.append(env.render(env_ids=[0, 1, 2, 3]).cpu())
    print(f"{args.steps} steps x {n} envs; mean team reward per step {float(ret.mean()) / args.steps:+.4f}; "
          f"frames {tuple(frames[0].shape)} uint8 x {len(frames)}")
    if args.out:
        np.save(args.out, torch.stack(frames).numpy())
    env.close()


if __name__ == "__main__":
    main()
