"""Per-rank e2e (host buffers, delta transport) under torchrun with several (batches in flight, host threads) settings."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("gloo")
n, RING, steps = 65536, 16, 300
rng = np.random.default_rng(rank)
acts = [rng.integers(0, 4, size=(RING, n, 2)).astype(np.int8) for _ in range(2)]
for EB, threads, transport in ((2, 4, "delta"), (2, 3, "delta"), (2, 2, "delta"), (1, 3, "delta"), (1, 2, "delta"), (1, 4, "delta"), (2, 3, "full")):
    envs = [mg.make_vec("multigrid-collect-respawn-clustered-v0", n, device=dev, seed=0, env_id_base=(rank * 2 + b) * n, host_threads=threads, host_transport=transport) for b in range(EB)]
    for e in envs:
        e.reset()
    for i in range(2 * EB + 2):
        envs[i % EB].step(acts[i % EB][i % RING])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for b in range(EB):
        envs[b].step_async(acts[b][0])
    for i in range(steps):
        b = i % EB
        out = envs[b].step_wait()
        if i + EB < steps:
            envs[b].step_async(acts[b][(i // EB + 1) % RING])
    dt = time.perf_counter() - t0
    rate = torch.tensor([n * steps / dt], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rate)
    if rank == 0:
        print(f"EB={EB} threads={threads} {transport}: total {float(rate):.3e} env-steps/s over {world} ranks ({float(rate)/world:.3e} per rank, {dt/steps*1e6:.0f} us per step on rank 0)", flush=True)
    for e in envs:
        e.close()
    if world > 1:
        dist.barrier()
