#!/usr/bin/env python
"""Host-side cost of one `step()` call (Python + ctypes + launch), measured on a batch small enough that the GPU is never
the bottleneck: wall time per call over many back-to-back calls (development aid)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_multigrid_b200 as mg  # noqa: E402


def main():
    n = 1024
    envs = mg.make_vec("multigrid-collect-respawn-clustered-v0", n)
    envs.reset()
    a = torch.randint(0, 4, (n, 2), device="cuda:0", dtype=torch.int8)
    for _ in range(200):
        envs.step(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 20000
    for _ in range(K):
        envs.step(a)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"CollectVecEnv.step: {(t1 - t0) / K * 1e6:.2f} us per call on the host (queue drained {1e3 * (t2 - t1):.2f} ms after the loop)")
    envs.close()


if __name__ == "__main__":
    main()
