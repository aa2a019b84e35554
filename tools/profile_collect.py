#!/usr/bin/env python
"""A few launches of the Collect step kernel over rotating batches (so that the captured launch reads its state from
HBM, not L2) for `ncu -k regex:collect_step -s 20 -c 1` (development aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg  # noqa: E402

n, B = 65536, 8
envs = [mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=0, env_id_base=b * n) for b in range(B)]
acts = [torch.randint(0, 4, (n, 2), device="cuda:0", dtype=torch.int8) for _ in range(B)]
for e in envs:
    e.reset()
for i in range(40):
    envs[i % B].step(acts[i % B])
torch.cuda.synchronize()
for e in envs:
    e.close()
