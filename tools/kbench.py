#!/usr/bin/env python
"""Kernel micro-benchmark / phase-timeline tool for the Collect step kernel (development aid).

    python tools/kbench.py [--tiles 0,1,2,...] [--num-envs 65536] [--batches 16] [--steps 2000] [--timeline]

For every tile variant (MG_TILE): CUDA-graph replay of the fused step kernel over rotating env
batches (same protocol as bench.py), the stand-alone encode kernel, and - for scale - a torch
device copy moving the same number of bytes.  --timeline dumps per-CTA phase timestamps of one launch.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg  # noqa: E402

ENV_ID = "multigrid-collect-respawn-clustered-v0"


def graph_time(fn_list, reps, stream):
    with torch.cuda.stream(stream):
        for f in fn_list:
            f()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for f in fn_list:
                f()
        g.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            g.replay()
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fn_list))  # us per call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", default="0")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--batches", type=int, default=16)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--autoreset", type=int, default=1)
    ap.add_argument("--time-limit", type=int, default=None, help="override max_episode_steps (1 = every step autoresets)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n, B = args.num_envs, args.batches
    stream = torch.cuda.Stream(device=dev)
    out = {"num_envs": n, "batches": B}

    # scale: a plain device copy of the same algorithmic bytes (592 B/env: 296 in, 296 out)
    src = [torch.empty(n * 296, dtype=torch.uint8, device=dev) for _ in range(B)]
    dst = [torch.empty(n * 296, dtype=torch.uint8, device=dev) for _ in range(B)]
    us = graph_time([lambda a=a, b=b: b.copy_(a) for a, b in zip(src, dst)], max(1, args.steps // B), stream)
    out["torch_copy_same_bytes_us"] = us
    out["torch_copy_GBps"] = n * 592 / us / 1e3
    del src, dst

    for tile in [int(t) for t in args.tiles.split(",")]:
        os.environ["MG_TILE"] = str(tile)
        extra = {} if args.time_limit is None else {"max_episode_steps": args.time_limit}
        envs = [mg.make_vec(ENV_ID, n, device=dev, seed=0, autoreset=bool(args.autoreset), env_id_base=b * n, **extra) for b in range(B)]
        gen = torch.Generator(device=dev).manual_seed(1)
        acts = [torch.randint(0, 4, (n, 2), generator=gen, device=dev, dtype=torch.int8) for _ in range(B)]
        for e in envs:
            e.reset()
        E = envs[0]._lib.mg_tile_envs(envs[0]._h)
        us = graph_time([lambda e=e, a=a: e.step(a) for e, a in zip(envs, acts)], max(1, args.steps // B), stream)
        us_enc = graph_time([lambda e=e: e.encode() for e in envs], max(1, args.steps // B), stream)
        r = {"tile": tile, "E": E, "step_us": us, "step_GBps_592": n * 592 / us / 1e3, "env_steps_per_s": n / us * 1e6,
             "encode_us": us_enc, "encode_GBps_400": n * 400 / us_enc / 1e3}
        if args.timeline:
            tiles = (n + E - 1) // E
            tl = torch.zeros((tiles, 8), dtype=torch.int64, device=dev)
            e = envs[0]
            e._lib.mg_debug_set_timeline(e._h, C.c_void_p(tl.data_ptr()))
            for b in range(1, B):   # evict batch 0 from L2
                envs[b].step(acts[b])
            e.step(acts[0])
            torch.cuda.synchronize()
            e._lib.mg_debug_set_timeline(e._h, None)
            t = tl.cpu().numpy().astype(np.int64)
            t0 = t[:, 0].min()
            rel = t[:, :6] - t0
            ph = np.diff(rel, axis=1)
            names = ["load_wait", "step", "autoreset+hdr", "expand", "store_drain"]
            r["timeline_ns"] = {
                "span": int(rel[:, 5].max()),
                "cta_start_p50_p90_max": [int(np.percentile(rel[:, 0], q)) for q in (50, 90, 100)],
                "cta_end_p10_p50_max": [int(np.percentile(rel[:, 5], q)) for q in (10, 50, 100)],
                **{nm: [int(np.percentile(ph[:, i], q)) for q in (10, 50, 90, 100)] for i, nm in enumerate(names)},
            }
        print(json.dumps(r), flush=True)
        out.setdefault("variants", []).append(r)
        for e in envs:
            e.close()
        del envs, acts
        torch.cuda.empty_cache()
    print(json.dumps({k: v for k, v in out.items() if k != "variants"}))


if __name__ == "__main__":
    main()
