#!/usr/bin/env python
"""A few launches of the Collect step over rotating env batches with fresh actions every step (so that the captured launch
reads its state from HBM, not L2, and takes the pickup / respawn paths of the benchmark) for
`ncu -k regex:"collect_rollout|collect_step" -s 20 -c 1` (development aid).  MG_STEP_IMPL=tile selects the CTA-tile kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg  # noqa: E402

n, B, RING = 65536, 8, 16
envs = [mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=0, env_id_base=b * n) for b in range(B)]
gen = torch.Generator(device="cuda:0").manual_seed(1)
acts = [torch.randint(0, 4, (RING, n, 2), generator=gen, device="cuda:0", dtype=torch.int8) for _ in range(B)]
for e in envs:
    e.reset()
torch.cuda.synchronize()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    envs[i % B].step(acts[i % B][(i // B) % RING])
torch.cuda.synchronize()
assert max(e.status() for e in envs) == 0
for e in envs:
    e.close()
