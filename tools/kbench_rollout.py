#!/usr/bin/env python
"""Development micro-benchmark of mg_rollout (T steps per launch, state resident in shared memory).

    python tools/kbench_rollout.py [--num-envs 4096,65536] [--steps 64] [--modes actions,policy,noobs]

Per configuration: B independent env batches (working set > L2 for the large sizes), each rolled out for T steps per launch with
its own fresh action tensor [T, N, A]; CUDA events around `reps` rounds over the batches.  Reported per env-step: microseconds per
step (launch time / T), env-steps/s, and GB/s of the bytes a rollout step actually moves (actions 2 + obs 300 + rewards 16 +
flags 2 = 320 B per env-step, + 2 x 136 B of state per env per LAUNCH) against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg  # noqa: E402

ENV_ID = "multigrid-collect-respawn-clustered-v0"


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0


def run(n, T, mode, reps=5):
    dev = torch.device("cuda:0")
    per_launch = n * (T * 320 + 272)
    B = max(2, min(8, int(300e6 // per_launch) + 1))
    envs = [mg.make_vec(ENV_ID, n, device=dev, seed=0, env_id_base=b * n) for b in range(B)]
    gen = torch.Generator(device=dev).manual_seed(1)
    acts = [torch.randint(0, 4, (T, n, 2), generator=gen, device=dev, dtype=torch.int8) for _ in range(B)]
    outs = []
    for e in envs:
        e.reset()
        o = dict(rewards=torch.empty((T, n, 2), dtype=torch.float64, device=dev), terminated=torch.empty((T, n), dtype=torch.uint8, device=dev),
                 truncated=torch.empty((T, n), dtype=torch.uint8, device=dev))
        if mode != "noobs":
            o["obs"] = torch.empty((T, n, 10, 10, 3), dtype=torch.uint8, device=dev)
        if mode == "policy":
            o["actions"] = torch.empty((T, n, 2), dtype=torch.int8, device=dev)
        outs.append(o)
    torch.cuda.synchronize(dev)

    def call(b):
        if mode == "policy":
            envs[b].rollout(steps=T, out=outs[b])
        else:
            envs[b].rollout(acts[b], obs=(mode != "noobs"), out=outs[b])

    main = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(main):
        for b in range(B):
            call(b)
        main.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            for b in range(B):
                call(b)
        g.replay()
        main.synchronize()
        samples = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            g.replay()
            e1.record(main)
            main.synchronize()
            samples.append(e0.elapsed_time(e1) * 1e3 / B)
    st = [e.status() for e in envs]
    for e in envs:
        e.close()
    assert max(st) == 0, st
    us = statistics.median(samples)
    bytes_step = (320 if mode != "noobs" else 20) + 272 / T
    return {"num_envs": n, "T": T, "mode": mode, "batches": B, "us_per_launch": round(us, 2), "us_per_step": round(us / T, 4),
            "env_steps_per_s": n * T / us * 1e6, "bytes_per_env_step_moved": round(bytes_step, 1),
            "GBps_moved": round(n * T * bytes_step / us / 1e3, 1), "frac_moved": round(n * T * bytes_step / us / 1e3 / peak(), 4),
            "frac_592B": round(n * T * 592 / us / 1e3 / peak(), 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", default="4096,65536")
    ap.add_argument("--steps", default="1,8,64")
    ap.add_argument("--modes", default="actions,policy,noobs")
    args = ap.parse_args()
    for n in [int(x) for x in args.num_envs.split(",")]:
        for T in [int(x) for x in args.steps.split(",")]:
            for mode in args.modes.split(","):
                print(json.dumps(run(n, T, mode)), flush=True)


if __name__ == "__main__":
    main()
