#!/usr/bin/env python
"""Random-policy rollout on one B200: the gymnasium-style loop a user of the reference would write, batched.

    python examples/random_rollout.py [--env-id multigrid-collect-respawn-clustered-v0] [--num-envs 65536] [--steps 500]

Everything stays on the device: actions are sampled with torch, `step` enqueues one fused kernel and returns views of
reused output tensors (no allocation, no host sync), rewards are accumulated on the GPU; the only synchronisation is the
final `.item()`.  With `--graph` the whole step (action sampling + env step + reward accumulation) is captured once in a
CUDA graph and replayed, which removes the per-step Python / launch overhead for small batches.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_multigrid_b200 as mg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env-id", default="multigrid-collect-respawn-clustered-v0")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--graph", action="store_true")
    args = ap.parse_args()

    envs = mg.make_vec(args.env_id, args.num_envs, device="cuda:0", seed=0)
    obs, info = envs.reset()
    n, A = envs.num_envs, envs.num_agents
    returns = torch.zeros(n, A, dtype=torch.float64, device="cuda:0")
    episodes = torch.zeros((), dtype=torch.int64, device="cuda:0")
    actions = torch.empty((n, A), dtype=torch.int8, device="cuda:0")

    def one_step():
        actions.random_(0, 4)                                   # Discrete(4): north / east / south / west
        obs, rewards, terminated, truncated, info = envs.step(actions)
        returns.add_(rewards)
        episodes.add_((terminated | truncated).sum())

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for _ in range(3):
            one_step()
        if args.graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                one_step()
            run = g.replay
        else:
            run = one_step
        stream.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run()
        stream.synchronize()
        dt = time.perf_counter() - t0
    print(f"{args.steps} steps x {n} envs in {dt * 1e3:.1f} ms: {args.steps * n / dt:.3e} env-steps/s, "
          f"{int(episodes.item())} episodes finished, mean reward per env-step {float(returns.sum()) / ((args.steps + 3) * n):.4f}")
    envs.close()


if __name__ == "__main__":
    main()
