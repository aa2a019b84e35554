import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg
n, EB, RING, steps = 65536, 2, 32, 400
for threads in (16, 8, 4, 2, 1):
    envs = [mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=0, env_id_base=b * n, host_threads=threads) for b in range(EB)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(0)
    acts = [rng.integers(0, 4, size=(RING, n, 2)).astype(np.int8) for _ in range(EB)]
    for i in range(6):
        envs[i % EB].step(acts[i % EB][i % RING])
    torch.cuda.synchronize()
    tw = ta = 0.0
    t0 = time.perf_counter()
    for b in range(EB):
        envs[b].step_async(acts[b][0])
    for i in range(steps):
        b = i % EB
        t1 = time.perf_counter()
        out = envs[b].step_wait()
        t2 = time.perf_counter()
        envs[b].step_async(acts[b][(i // EB + 1) % RING])
        t3 = time.perf_counter()
        tw += t2 - t1; ta += t3 - t2
    for b in range(EB):
        envs[b].step_wait()
    dt = time.perf_counter() - t0
    print(f"threads={threads:2d}: {n*steps/dt:.3e} env-steps/s, per step {dt/steps*1e6:.1f} us = wait+decode {tw/steps*1e6:.1f} + async enqueue {ta/steps*1e6:.1f}", flush=True)
    for e in envs:
        e.close()
