import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg
n, T = int(sys.argv[1]), int(sys.argv[2])
obs = (sys.argv[3] != "noobs") if len(sys.argv) > 3 else True
env = mg.make_vec("multigrid-collect-respawn-clustered-v0", n, device="cuda:0", seed=0)
env.reset()
act = torch.randint(0, 4, (T, n, 2), device="cuda:0", dtype=torch.int8)
for _ in range(3):
    env.rollout(act, obs=obs)
tiles = (n + 3) // 4      # rows for the smallest tile size (4 envs per warp); unused rows stay zero
tl = torch.zeros((tiles, 8), dtype=torch.int64, device="cuda:0")
env._lib.mg_debug_set_timeline(env._h, C.c_void_p(tl.data_ptr()))
env.rollout(act, obs=obs)
torch.cuda.synchronize()
env._lib.mg_debug_set_timeline(env._h, None)
t = tl.cpu().numpy().astype(np.float64)
t = t[t[:, 7] > 0]
tiles = len(t)
names = ["actions", "step", "wait_read", "encode+delta", "reset", "store", "total", "T"]
print(f"n={n} T={T} obs={obs}: cycles per step (median over {tiles} tiles; total includes the state load / store)")
for i, k in enumerate(names[:7]):
    print(f"  {k:14s} median {np.median(t[:, i]) / T:9.0f}   max {np.max(t[:, i]) / T:9.0f}")
