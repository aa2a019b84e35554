import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg
n, T = int(sys.argv[1]), int(sys.argv[2])
env = mg.make_vec("multigrid-collect-respawn-clustered-v0", n, device="cuda:0", seed=0)
env.reset()
act = torch.randint(0, 4, (T, n, 2), device="cuda:0", dtype=torch.int8)
for _ in range(4):
    env.rollout(act)
torch.cuda.synchronize()
assert env.status() == 0
print("ok")
