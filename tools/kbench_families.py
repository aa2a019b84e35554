#!/usr/bin/env python
"""Micro-benchmark of every kernel family besides the Collect step (development aid; tools/kbench.py covers Collect).

    python tools/kbench_families.py [--which ctf,maze,view,toroid,wildfire,generic,render,collect_streams] [--reps 40]

Same protocol as bench.py: CUDA graph over B independent env batches whose working set exceeds the 126 MB L2,
CUDA events on the launching stream, warm-up first.  Each line reports the algorithmic bytes per env-step of that
configuration and the achieved GB/s against MEASURED_PEAKS.json.  Runs on the GPU box only:
maps come from the committed golden fixtures, nothing under /root/reference is read.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0


def golden(stem, key):
    with np.load(os.path.join(GOLDEN, stem + ".npz")) as z:
        return z[key]


def graph_time(fns, reps, streams=1):
    """us per call of the callables in `fns`, captured once into a CUDA graph (forked over `streams` streams)."""
    dev = torch.device("cuda:0")
    torch.cuda.synchronize(dev)      # resets / setup ran on the default stream; the streams below are non-blocking
    main = torch.cuda.Stream(device=dev)
    side = [torch.cuda.Stream(device=dev) for _ in range(streams - 1)]
    with torch.cuda.stream(main):
        for f in fns:
            f()
        main.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            for s in side:
                s.wait_stream(main)
            for i, f in enumerate(fns):
                st = main if i % streams == 0 else side[i % streams - 1]
                with torch.cuda.stream(st):
                    f()
            for s in side:
                main.wait_stream(s)
        for _ in range(3):
            g.replay()
        main.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _ in range(reps):
            g.replay()
        e1.record(main)
        main.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))


RING = 8


def ring_steps(envs, draw):
    """Callables for one CUDA graph: launch g steps env batch g % B with ring slot (g // B) % RING of `draw(b)`-shaped fresh
    actions - every step of every batch gets its own independently drawn action tensor (a constant tensor parks the agents
    against a wall after a few steps, which hides most of the dynamics: pickups, battles, moving fire fighters)."""
    rings = [[draw(b) for _ in range(RING)] for b in range(len(envs))]
    B = len(envs)
    return [lambda e=envs[g % B], a=rings[g % B][(g // B) % RING]: e.step(a) for g in range(B * RING)]


def report(name, n, us, bytes_per_env, **extra):
    gbs = n * bytes_per_env / us / 1e3
    print(json.dumps({"kernel": name, "num_envs": n, "us_per_launch": round(us, 3), "algorithmic_bytes_per_env": bytes_per_env,
                      "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak(), 4), "env_steps_per_s": n / us * 1e6, **extra}),
          flush=True)


def batches_for(bytes_per_env, n, want=400e6, lo=2, hi=32):
    return int(max(lo, min(hi, np.ceil(want / (bytes_per_env * n)))))


def bench_ctf(args):
    fm = golden("ctf_2v2", "field_map")
    for nb, nr, n in ((2, 2, 65536), (2, 2, 1 << 20), (8, 8, 1 << 18)):
        if nb + nr > 8 and fm.shape[0] < 10:
            continue
        bpe = 100 + 2 * (4 * (nb + nr) + 16) + nb + 10          # obs u8 + state r/w + actions + reward/flags
        B = batches_for(bpe, n)
        envs = [mg.make_ctf_vec(n, fm, num_blue_agents=nb, num_red_agents=nr, seed=b, env_id_base=b * n) for b in range(B)]
        for e in envs:
            e.reset()
        torch.cuda.synchronize()
        us = graph_time(ring_steps(envs, lambda b: torch.randint(0, 5, (n, nb), device="cuda:0", dtype=torch.int8)), max(2, args.reps // RING))
        report(f"map_kernel<ctf> {nb}v{nr} 10x10 map obs u8", n, us, bpe, batches=B)
        for e in envs:
            e.close()
    n = 262144
    e = mg.make_ctf_vec(n, fm, seed=0)
    e.reset()
    outs = [torch.empty((n, 216), dtype=torch.int64, device="cuda:0") for _ in range(2)]
    us = graph_time([lambda o=o: e.flattened_obs(out=o) for o in outs], args.reps)
    report("ctf_flat_kernel 2v2 flattened obs int64 [216]", n, us, 216 * 8 + 16, batches=2)
    n8 = 1 << 20
    e8 = mg.make_ctf_vec(n8, fm, seed=0)
    e8.reset()
    outs8 = [torch.empty((n8, 216), dtype=torch.uint8, device="cuda:0") for _ in range(2)]
    us = graph_time([lambda o=o: e8.flattened_obs(out=o) for o in outs8], args.reps)
    report("ctf_flat_kernel 2v2 flattened obs uint8 [216]", n8, us, 216 + 16, batches=2)
    e8.close()
    e.close()
    n = 65536
    envs = [mg.make_ctf_vec(n, fm, reference_dtypes=True, seed=b, env_id_base=b * n) for b in range(8)]
    for e in envs:
        e.reset()
    torch.cuda.synchronize()
    us = graph_time(ring_steps(envs, lambda b: torch.randint(0, 5, (n, 2), device="cuda:0", dtype=torch.int8)), max(2, args.reps // RING))
    report("map_kernel<ctf> 2v2 10x10 map obs int64 (reference dtype)", n, us, 800 + 2 * 32 + 12, batches=8)
    for e in envs:
        e.close()


def bench_ctf_policy(args):
    """CtF 2v2 with the scripted opponents decided on the device: the policy launch alone, and policy + step per env-step."""
    from gym_multigrid_b200.policy.ctf.heuristic import FightPolicy, PatrolFightPolicy
    fm = golden("ctf_2v2", "field_map")
    fmf = fm.astype(np.float64)
    n, nb, nr = 1 << 20, 2, 2
    bpe_step = 100 + 2 * (4 * (nb + nr) + 16) + nb + 10
    bpe_pol = 4 * (nb + nr) + 16 + nr                        # agents row + header read, red actions written
    B = batches_for(bpe_step, n)
    envs = [mg.make_ctf_vec(n, fm, num_blue_agents=nb, num_red_agents=nr, seed=b, env_id_base=b * n) for b in range(B)]
    for e in envs:
        e.set_enemy_policies([FightPolicy(fmf), PatrolFightPolicy(fmf)], device=True)
        e.reset()
    torch.cuda.synchronize()
    us = graph_time(ring_steps(envs, lambda b: torch.randint(0, 5, (n, nb), device="cuda:0", dtype=torch.int8)), max(2, args.reps // RING))
    report("map_kernel<ctf> 2v2 with the fused policy prologue (fight, patrol_fight reds decided on the device)", n, us, bpe_step + bpe_pol, batches=B)
    import ctypes as C
    ptr = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    us = graph_time([lambda e=e: e._check(e._lib.mg_red_policy_actions(e._h, ptr(e.state), ptr(e._red_buf), e._stream())) for e in envs], args.reps)
    report("ctf_policy_kernel alone 2v2", n, us, bpe_pol, batches=B)
    for e in envs:
        e.close()


def bench_maze(args):
    fm = golden("maze_gen64", "field_map")
    for n, ref in ((16384, False), (131072, False), (16384, True)):
        elem = 8 if ref else 1
        bpe = 4096 * elem + 2 * (4 + 16) + 1 + 10
        B = batches_for(bpe, n)
        envs = [mg.make_maze_vec(n, fm, seed=b, env_id_base=b * n, reference_dtypes=ref) for b in range(B)]
        draw = lambda b: torch.randint(0, 5, (n,), device="cuda:0", dtype=torch.int8)  # noqa: E731
        for e in envs:
            e.reset()
        torch.cuda.synchronize()
        us = graph_time(ring_steps(envs, draw), max(2, args.reps // RING))
        report(f"map_kernel<maze> 64x64 full-map obs {'float64 (reference dtype)' if ref else 'u8'}", n, us, bpe, batches=B)
        if not ref:
            outs = [torch.empty((n, 1, 7, 7, 3), dtype=torch.uint8, device="cuda:0") for _ in range(B)]
            us = graph_time([lambda e=e, o=o: e.gen_obs(7, False, out=o) for e, o in zip(envs, outs)], args.reps)
            report("view_kernel maze 64x64 V=7 partial obs", n, us, 147 + 4, batches=B)
            for e in envs:
                e.set_partial_obs(7)
            us = graph_time(ring_steps(envs, draw), max(2, args.reps // RING))
            report("map_kernel<maze> 64x64 fused step + V=7 partial obs (config 4)", n, us, 147 + 2 * (4 + 16) + 1 + 10, batches=B)
        for e in envs:
            e.close()


def bench_maze_partial(args):
    """Config 4 (fused step + V = 7 partial views) over env counts: one wave (131 072) up to many waves (1 M)."""
    fm = golden("maze_gen64", "field_map")
    for n in (131072, 524288, 1 << 20):
        B = 3 if n >= 524288 else 8
        envs = [mg.make_maze_vec(n, fm, seed=b, env_id_base=b * n) for b in range(B)]
        for e in envs:
            e.set_partial_obs(7)
            e.reset()
        torch.cuda.synchronize()
        us = graph_time(ring_steps(envs, lambda b: torch.randint(0, 5, (n,), device="cuda:0", dtype=torch.int8)), max(2, args.reps // RING))
        report("map_kernel<maze> 64x64 fused step + V=7 partial obs (config 4)", n, us, 147 + 2 * (4 + 16) + 1 + 10, batches=B)
        for e in envs:
            e.close()
        del envs
        torch.cuda.empty_cache()


def bench_view(args):
    for env_id, V, n in (("multigrid-collect-respawn-clustered-v0", 7, 65536), ("multigrid-collect-rooms-respawn-v0", 5, 65536),
                         ("multigrid-collect-respawn-clustered-v0", 7, 1 << 20)):
        B = 8 if n <= 65536 else 2
        envs = [mg.make_vec(env_id, n, seed=b, env_id_base=b * n) for b in range(B)]
        for e in envs:
            e.reset()
        A, cells = envs[0].num_agents, envs[0].width * envs[0].height
        outs = [torch.empty((n, A, V, V, 3), dtype=torch.uint8, device="cuda:0") for _ in range(B)]
        us = graph_time([lambda e=e, o=o: e.gen_obs(V, False, out=o) for e, o in zip(envs, outs)], args.reps)
        report(f"view_kernel collect {env_id} V={V}", n, us, cells + 2 * A + A * V * V * 3, batches=B)
        touts = [torch.empty((n, A, envs[0].width, envs[0].height, envs[0].num_ball_types + A), dtype=torch.float32, device="cuda:0")
                 for _ in range(B)]
        us = graph_time([lambda e=e, o=o: e.toroid_obs(out=o) for e, o in zip(envs, touts)], args.reps)
        report(f"toroid_kernel {env_id}", n, us, cells + 2 * A + touts[0][0].numel() * 4, batches=B)
        for e in envs:
            e.close()


def bench_wildfire(args):
    for size, A, n in ((64, 16, 16384), (64, 16, 131072), (64, 32, 131072), (128, 32, 32768), (32, 32, 65536)):
        cells = size * size
        bpe = A + 2 * (cells + 4 * A + 16) + 3 * cells + 8 * A + 2
        B = batches_for(bpe, n, lo=2, hi=8)
        envs = [mg.make_wildfire_vec(n, size=size, num_agents=A, seed=b, env_id_base=b * n) for b in range(B)]
        for e in envs:
            e.reset()
        torch.cuda.synchronize()
        steps = ring_steps(envs, lambda b: torch.randint(0, 5, (n, A), device="cuda:0", dtype=torch.int8))
        for _ in range(4):           # let the fires develop: the stencil skips quiet groups
            for f in steps:
                f()
        us = graph_time(steps, max(2, args.reps // (4 * RING)))
        report(f"wildfire_kernel {size}x{size} A={A}", n, us, bpe, batches=B)
        for e in envs:
            e.close()


def bench_memset(args):
    """Write-only and copy ceilings measured with the same protocol (graph over rotating buffers > L2)."""
    nbytes = 256 << 20
    bufs = [torch.empty(nbytes, dtype=torch.uint8, device="cuda:0") for _ in range(4)]
    us = graph_time([lambda b=b: b.zero_() for b in bufs], args.reps)
    print(json.dumps({"kernel": "torch zero_ (write only)", "bytes": nbytes, "us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1),
                      "frac_of_measured_peak": round(nbytes / us / 1e3 / peak(), 4)}), flush=True)
    us = graph_time([lambda a=bufs[i], b=bufs[(i + 1) % 4]: b.copy_(a) for i in range(4)], args.reps)
    print(json.dumps({"kernel": "torch copy_ (read + write)", "bytes": 2 * nbytes, "us": round(us, 2), "GBps": round(2 * nbytes / us / 1e3, 1),
                      "frac_of_measured_peak": round(2 * nbytes / us / 1e3 / peak(), 4)}), flush=True)
    x = [torch.empty(nbytes // 4, dtype=torch.float32, device="cuda:0") for _ in range(4)]
    us = graph_time([lambda b=b: b.sum() for b in x], args.reps)
    print(json.dumps({"kernel": "torch sum (read only)", "bytes": nbytes, "us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1),
                      "frac_of_measured_peak": round(nbytes / us / 1e3 / peak(), 4)}), flush=True)


def bench_generic(args):
    g = {k: golden("generic_12x12_a5", k) for k in ("init_obs", "init_pos")}
    n, A, S = 65536, 5, 12
    cells = S * S
    bpe = A + 2 * (2 * cells + 2 * A + 16) + A * cells * 6 + 8 * A + 2
    B = 4
    envs = []
    for b in range(B):
        e = mg.make_generic_vec(n, S, num_agents=A, max_steps=60, seed=b, env_id_base=b * n)
        idx = np.arange(n) % g["init_obs"].shape[0]
        e.set_layout(g["init_obs"][idx, 0], g["init_pos"][idx])
        e.reset()
        envs.append(e)
    torch.cuda.synchronize()
    us = graph_time(ring_steps(envs, lambda b: torch.randint(0, 4, (n, A), device="cuda:0", dtype=torch.int8)), max(2, args.reps // RING))
    report("generic_kernel 12x12 A=5 encode_dim 6 obs per agent", n, us, bpe, batches=B)
    for e in envs:
        e.close()


def bench_render(args):
    """render(): frames of the first n envs, tile sizes 32 (the reference default) and 8; bytes = the frame + the cells read."""
    N = 4096
    env = mg.make_vec("multigrid-collect-respawn-clustered-v0", N, seed=0)
    env.reset()
    for ts, n in ((32, 1024), (32, 64), (8, 4096)):
        frame = 10 * ts * 10 * ts * 3
        B = max(2, int(np.ceil(400e6 / (frame * n))))
        outs = [torch.empty((n, 10 * ts, 10 * ts, 3), dtype=torch.uint8, device="cuda:0") for _ in range(B)]
        ids = torch.arange(n, device="cuda:0", dtype=torch.int32)
        us = graph_time([lambda o=o: env.render(env_ids=ids, tile_size=ts, out=o) for o in outs], args.reps)
        report(f"render_kernel Collect 10x10 tile_size {ts}", n, us, frame + 100, batches=B)
    env.close()
    fm = golden("maze_gen64", "field_map")
    m = mg.make_maze_vec(256, fm)
    m.reset()
    for ts, n in ((8, 256),):
        frame = 64 * ts * 64 * ts * 3
        outs = [torch.empty((n, 64 * ts, 64 * ts, 3), dtype=torch.uint8, device="cuda:0") for _ in range(2)]
        us = graph_time([lambda o=o: m.render(tile_size=ts, out=o) for o in outs], args.reps)
        report(f"render_kernel Maze 64x64 tile_size {ts}", n, us, frame + 4, batches=2)
    m.close()


def bench_collect_streams(args):
    """Collect step: the 16 independent env batches forked over 1/2/4 streams inside the graph."""
    n, B = 65536, 16
    envs = [mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=0, env_id_base=b * n) for b in range(B)]
    acts = [torch.randint(0, 4, (n, 2), device="cuda:0", dtype=torch.int8) for _ in range(B)]
    for e in envs:
        e.reset()
    for S in (1, 2, 4, 8):
        us = graph_time([lambda e=e, a=a: e.step(a) for e, a in zip(envs, acts)], args.reps * 4, streams=S)
        report(f"collect_step_kernel streams={S}", n, us, 592, batches=B)
    for e in envs:
        e.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="ctf,maze,view,wildfire,generic")
    ap.add_argument("--reps", type=int, default=40)
    args = ap.parse_args()
    for w in args.which.split(","):
        globals()["bench_" + w](args)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
