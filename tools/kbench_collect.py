#!/usr/bin/env python
"""Development micro-benchmark of the Collect step kernel under the bench's protocol (16 independent env batches, a ring of
64 fresh action tensors per batch, CUDA graphs, CUDA events), swept over kernel variants selected by environment switches
that mg_create reads (MG_TILE, MG_EARLY_OBS, MG_STEP_IMPL ...).

    python tools/kbench_collect.py [--num-envs 65536] [--variants "MG_TILE=0;MG_TILE=0,MG_EARLY_OBS=0;..."]
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


def run(n, B, RING, streams, launches, reps, constant=False):
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1)
    envs = [mg.make_vec(ENV_ID, n, device=dev, seed=0, autoreset=True, env_id_base=b * n) for b in range(B)]
    rings = [torch.randint(0, 4, (RING, n, 2), generator=gen, device=dev, dtype=torch.int8) for _ in range(B)]
    for e in envs:
        e.reset()
    torch.cuda.synchronize(dev)      # reset ran on the default stream; `main` below is a non-blocking stream
    main = torch.cuda.Stream(device=dev)
    side = [torch.cuda.Stream(device=dev) for _ in range(streams - 1)]
    with torch.cuda.stream(main):
        for i in range(B):
            envs[i].step(rings[i][0])
        main.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            for s in side:
                s.wait_stream(main)
            for gi in range(launches):
                b = gi % B
                st = main if b % streams == 0 else side[b % streams - 1]
                with torch.cuda.stream(st):
                    envs[b].step(rings[b][0 if constant else (gi // B) % RING])
            for s in side:
                main.wait_stream(s)
        for _ in range(3):
            g.replay()
        main.synchronize()
        out = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            g.replay()
            e1.record(main)
            main.synchronize()
            out.append(e0.elapsed_time(e1) * 1e3 / launches)
    st = [e.status() for e in envs]
    for e in envs:
        e.close()
    if max(st):
        print(json.dumps({"WARNING": "device status words", "status": st}), flush=True)
    return statistics.median(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", default="65536,4096")
    ap.add_argument("--variants", default="MG_TILE=0")
    ap.add_argument("--streams", default="1,4")
    ap.add_argument("--constant", action="store_true", help="also time with one constant action tensor (round 1's degenerate workload)")
    args = ap.parse_args()
    for var in args.variants.split(";"):
        kv = dict(x.split("=") for x in var.split(",") if x)
        old = {k: os.environ.get(k) for k in kv}
        os.environ.update(kv)
        for n in [int(x) for x in args.num_envs.split(",")]:
            for S in [int(x) for x in args.streams.split(",")]:
                for const in ([False, True] if args.constant else [False]):
                    us = run(n, 16, 64, S, 1024, 5, const)
                    print(json.dumps({"variant": var, "num_envs": n, "streams": S, "constant_actions": const, "us_per_launch": round(us, 3),
                                      "frac_592B": round(592 * n / us / 1e3 / peak(), 4), "env_steps_per_s": n / us * 1e6}), flush=True)
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


if __name__ == "__main__":
    main()
