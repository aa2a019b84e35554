#!/usr/bin/env python
"""Benchmark of the hot path: batched `CollectGameEnv.step` + `Grid.encode` (BASELINE.json configs[1]:
multigrid-collect-respawn-clustered-v0, 2 agents, 65 536 envs per launch per GPU), with the other BASELINE configs
(CtF, Maze + partial views, Wildfire) timed beside it under `families`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one fused step+encode launch over one batch of 65 536 envs.  To defeat the 126 MB L2 the timed loop rotates
over `--batches` independent env batches (each with its own state, obs and action tensors; 16 x 39 MB = 630 MB working
set), so every launch streams its state from HBM.  ACTIONS ARE FRESH EVERY STEP: each env batch owns a ring of
`--action-ring` (64) independently drawn uniform action tensors and launch g of the run steps batch g % B with ring slot
(g // B) % 64 - on the device arm (inside the captured graphs), on the host-buffer arm and on the CPU arm alike.
The K-step region is timed `--repeats` times (each repeat continues the action sequence); the median is reported.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions used here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "multigrid-collect-respawn-clustered-v0"
ALGO_BYTES_PER_ENV_STEP = 592   # actions 2 + state 2x136 + obs 300 + f64 rewards 16 + flags 2
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def workload_config(args):
    """The `config` object - identical in both arms (ours / reference): it names the workload, not the implementation."""
    return {"workload": f"{ENV_ID}, 2 agents, {args.num_envs} envs per step per GPU, TimeLimit 50 + same-step autoreset, Philox RNG",
            "actions": f"fresh per step (ring of {args.action_ring} uniform random action tensors per env batch)",
            "num_envs_per_gpu_per_step": args.num_envs, "env_batches_per_gpu": args.batches, "action_ring": args.action_ring,
            "l2": f"inputs larger than L2: the timed loop rotates over {args.batches} independent env batches "
                  f"({args.batches * args.num_envs * (ALGO_BYTES_PER_ENV_STEP + 8) / 1e6:.0f} MB working set > 126 MB L2)",
            "repeats": args.repeats}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("collect_step_kernel_bytes_per_launch")
        except Exception:  # noqa: BLE001
            return None
    return None


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons through NVML: one synchronous sample at every timing event (`sample()`), plus a polling
    thread for regions long enough to be seen by it."""
    NAMES = None

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.active, self.sm, self.reasons, self.max_mhz = index, False, False, [], set(), None
        self.nv = self.h = None
        try:    # NVML is initialised HERE (caller's thread, before the timed region)
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
            }
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def sample(self):
        if self.nv is None:
            return
        try:
            self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.get_reasons(self.h)
            for bit, name in self.names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        while not self.stop_flag and self.nv is not None:
            if self.active:
                self.sample()
            time.sleep(0.001)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_env(num_envs, nthreads, seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as oc
    import gym_multigrid_b200 as mg
    s = mg.spec(ENV_ID)
    cfg = oc.make_collect_cfg(layout="quadrants_respawn", time_limit=s.max_episode_steps, **s.kwargs)
    o = oc.CollectOracle(cfg, num_envs, nthreads=nthreads)
    r = oc.PhiloxRng(seed=seed)
    o.reset(r)
    return o, r


def time_oracle(num_envs, steps, warmup, nthreads, ring=64, seed=0):
    """CPU restatement of the reference's path (oracle/mg_oracle.c, OpenMP over envs), same config, same kind of actions:
    a ring of `ring` independently drawn action arrays, a fresh one every step.  Returns (env-steps/s, seconds, pickups/env-step)."""
    import numpy as np
    o, r = oracle_env(num_envs, nthreads, seed)
    acts = np.random.default_rng(seed).integers(0, 4, size=(ring, num_envs, 2)).astype(np.int8)
    for i in range(warmup):
        o.step(acts[i % ring], r, autoreset=True, reuse_buffers=True)
    t0 = time.perf_counter()
    for i in range(steps):
        o.step(acts[(warmup + i) % ring], r, autoreset=True, reuse_buffers=True)
    dt = time.perf_counter() - t0
    # pickups per env-step under these actions (untimed), over two whole 50-step episodes like the GPU arm's count
    picked, PS = 0.0, 100
    for i in range(PS):
        picked += float(o.step(acts[(warmup + steps + i) % ring], r, autoreset=True, reuse_buffers=True)[1].sum())
    return num_envs * steps / dt, dt, picked / (PS * num_envs)


def python_reference(seconds=3.0):
    """The UNMODIFIED Python reference stepped on this box's host cores (single process / independent processes / one worker
    per core in lockstep over pipes = the AsyncVectorEnv pattern), when a copy of the package is present: /root/reference in
    the build container, baseline/_ref (oracle/install_reference.py) on the GPU box.  Runs in a child process."""
    for root in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(root, "gym_multigrid")):
            break
    else:
        return {"unavailable": "no copy of the reference package on this box (neither /root/reference nor baseline/_ref)"}
    env = dict(os.environ, MG_REFERENCE_ROOT=root)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS"):
        env.pop(k, None)
    try:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_python_baseline.py"), "--seconds", str(seconds), "--no-families"],
                             capture_output=True, text=True, timeout=120, env=env)
        out = json.loads(res.stdout.strip().splitlines()[-1])
        out["source"] = root
        return out
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = host_threads()
    # bounded sample: keep EXACTLY --steps steps, size the envs per step so the run ends in ~1 minute
    probe, _, _ = time_oracle(8192, 20, 2, nthreads)
    n = int(min(args.num_envs, max(512, probe * args.budget_s / max(1, args.steps + args.warmup))))
    n = max(64, n // 64 * 64)
    value, dt, pickups = time_oracle(n, args.steps, args.warmup, nthreads, ring=args.action_ring)
    sample = (f"{args.steps} steps x {n} envs per step ({dt:.1f} s; the GPU arm steps {args.num_envs} envs per launch), fresh actions per step, "
              "autoreset, Philox RNG, oracle/mg_oracle.c with OpenMP over envs")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "agent_steps_per_sec": value * 2,
        "config": workload_config(args),
        "pickups_per_env_step": pickups,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample, "num_envs_sampled": n,
                         "note": "CPU restatement (C, oracle/mg_oracle.c) of the reference's Python path on all host cores; the Python "
                                 "reference itself is quoted under python_reference when a copy of the package is on the box",
                         "python_reference": python_reference(args.python_seconds)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import gym_multigrid_b200 as mg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n, B, RING, R = args.num_envs, args.batches, args.action_ring, max(1, args.repeats)
    K, Wm = max(1, args.steps), max(3, args.warmup)
    PERIOD = B * RING                       # launches after which the (batch, ring slot) sequence repeats
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    envs, rings = [], []
    for b in range(B):
        e = mg.make_vec(ENV_ID, n, device=dev, seed=args.seed, autoreset=True, env_id_base=(rank * B + b) * n)
        e.reset()
        envs.append(e)
        rings.append(torch.randint(0, 4, (RING, n, 2), generator=gen, device=dev, dtype=torch.int8))
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput: CUDA graphs of launches, replayed.  Launch g of the run steps env batch g % B with the
    #      actions in ring slot (g // B) % RING.  The B env batches are independent, so a graph forks them over `--streams`
    #      streams (launches of one batch stay in order on one stream): tiles of one launch load while tiles of another compute
    #      / drain.  `single_stream` is the same sequence captured on ONE stream (whole launches serialised).
    graph_cache = {}

    def capture(n_streams, main, start, count):
        key = (n_streams, start % PERIOD, count)
        if key in graph_cache:
            return graph_cache[key]
        side = side_streams[:n_streams - 1]
        with torch.cuda.stream(main):
            main.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                for s_ in side:
                    s_.wait_stream(main)
                for i in range(count):
                    gi = start + i
                    b = gi % B
                    st = main if b % n_streams == 0 else side[b % n_streams - 1]
                    with torch.cuda.stream(st):
                        envs[b].step(rings[b][(gi // B) % RING])
                for s_ in side:
                    main.wait_stream(s_)
            g.replay()                         # a graph's first launch uploads it: keep that out of every timed region
            main.synchronize()
        graph_cache[key] = g
        return g

    CHUNK = max(PERIOD, 4096 - 4096 % PERIOD) if PERIOD <= 4096 else PERIOD

    def plan(n_streams, main, start, count):
        """Graphs whose replay, in order, performs launches [start, start + count) of the run."""
        out = []
        while count > 0:
            c = min(count, CHUNK)
            out.append(capture(n_streams, main, start, c))
            start, count = start + c, count - c
        return out

    # Every timed region starts behind an UNTIMED fill of a 160 MB scratch buffer on the same stream (the do_bench pattern): it writes
    # more than the 126 MB L2 holds, so no region starts with a warm cache, and the host enqueues the region's graphs while the fill
    # runs, so the device-side interval ev0 -> ev1 holds the K launches and not the host's graph-launch latency (with the driver's
    # --steps 20 a region is ~140 us: that latency was ~5 % of it; with the default K it is noise).
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)

    def timed_region(main, graphs):
        barrier()
        with torch.cuda.stream(main):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sampler.sample()
            sampler.active = True
            flush.zero_()
            ev0.record(main)
            for g in graphs:
                g.replay()
            ev1.record(main)
            main.synchronize()
            sampler.active = False
            sampler.sample()
        torch.cuda.synchronize(dev)
        return ev0.elapsed_time(ev1)

    def measure(n_streams, steps, repeats):
        main = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(main):
            for i in range(Wm):                # the W warm-up steps, eagerly (also allocates nothing later)
                envs[i % B].step(rings[i % B][(i // B) % RING])
            main.synchronize()
        plans = [plan(n_streams, main, Wm + r * steps, steps) for r in range(repeats)]
        warm = capture(n_streams, main, 0, PERIOD)
        with torch.cuda.stream(main):
            for _ in range(max(3, -(-2048 // PERIOD))):       # >= 2048 more untimed launches with fresh actions: settled clocks
                warm.replay()
            main.synchronize()
        return [timed_region(main, p) for p in plans]

    S = max(1, min(args.streams, B))
    side_streams = [torch.cuda.Stream(device=dev) for _ in range(S - 1)]
    n1 = max(B, min(K, 1024))
    single_ms = measure(1, n1, min(R, 3))
    single_us = statistics.median(single_ms) * 1e3 / n1
    all_ms = measure(S, K, R)
    ms = statistics.median(all_ms)
    status = max(e.status() for e in envs)
    assert status == 0, f"device status word {status}"

    # pickups per env-step under these actions (untimed): rewards are 1 per ball in this config
    acc = torch.zeros((), dtype=torch.float64, device=dev)
    PS = 100
    for i in range(PS):
        acc += envs[0].step(rings[0][i % RING])[1].sum()
    pickups = float(acc) / (PS * n)

    # ---- small batches: 4 096 envs per launch (the low end of BASELINE config 2), one stream
    small = None
    if not args.skip_small:
        ns = 4096
        se = [mg.make_vec(ENV_ID, ns, device=dev, seed=args.seed + 1, autoreset=True, env_id_base=(world * B + rank) * n + b * ns) for b in range(B)]
        sr = [torch.randint(0, 4, (RING, ns, 2), generator=gen, device=dev, dtype=torch.int8) for _ in range(B)]
        for e in se:
            e.reset()
        torch.cuda.synchronize(dev)     # the resets ran on the default stream; the streams below are non-blocking
        sm_main = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(sm_main):
            for i in range(B):
                se[i].step(sr[i][0])
            sm_main.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=sm_main):
                for gi in range(PERIOD):
                    se[gi % B].step(sr[gi % B][(gi // B) % RING])
            for _ in range(3):
                g.replay()
            sm_main.synchronize()
        t_small = statistics.median(timed_region(sm_main, [g]) for _ in range(3)) * 1e3 / PERIOD
        small = {"num_envs": ns, "avg_launch_us": t_small, "value": ns / t_small * 1e6,
                 "frac": ALGO_BYTES_PER_ENV_STEP * ns / t_small / 1e3 / measured_peak()[0],
                 "note": "one launch per step on ONE stream, CUDA graph, fresh actions"}
        for e in se:
            e.close()
        del se, sr

    # ---- launch-amortised rollouts (mg_rollout): T steps per launch with the env state held in shared memory, ONE stream.
    #      The loop `for t in range(T): env.step(actions[t])`; per env-step it moves actions 2 + obs 300 + rewards 16 + flags 2 =
    #      320 B, and 2 x 136 B of state per env once per LAUNCH - so against SURVEY's 592 B per env-step its fraction can exceed 1.
    rollout = None
    if not args.skip_rollout:
        rollout = {"T": args.rollout_steps, "bytes_moved_per_env_step": 320 + 272 / args.rollout_steps,
                   "note": "one launch = T steps of one env batch, state resident in shared memory; `policy` = uniform random actions drawn "
                           "on the device (Philox), `actions` = a given [T, N, A] tensor; one stream, CUDA events, median of 3"}
        Tr = args.rollout_steps
        for nn in (n, 4096):
            per_launch = nn * (Tr * 320 + 272)
            Br = max(1, min(4, int(300e6 // per_launch) + 1))
            re_ = [mg.make_vec(ENV_ID, nn, device=dev, seed=args.seed + 2, autoreset=True, env_id_base=((2 * world + rank) * B + b) * n) for b in range(Br)]
            outs = []
            for e in re_:
                e.reset()
                outs.append(dict(rewards=torch.empty((Tr, nn, 2), dtype=torch.float64, device=dev), terminated=torch.empty((Tr, nn), dtype=torch.uint8, device=dev),
                                 truncated=torch.empty((Tr, nn), dtype=torch.uint8, device=dev), obs=torch.empty((Tr, nn, 10, 10, 3), dtype=torch.uint8, device=dev),
                                 actions=torch.empty((Tr, nn, 2), dtype=torch.int8, device=dev)))
            racts = [torch.randint(0, 4, (Tr, nn, 2), generator=gen, device=dev, dtype=torch.int8) for _ in range(Br)]
            torch.cuda.synchronize(dev)
            for mode in ("policy", "actions"):
                rm = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(rm):
                    def call(b):
                        if mode == "policy":
                            re_[b].rollout(steps=Tr, out=outs[b])
                        else:
                            re_[b].rollout(racts[b], out=outs[b])
                    for b in range(Br):
                        call(b)
                    rm.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=rm):
                        for b in range(Br):
                            call(b)
                    g.replay()
                    rm.synchronize()
                us = statistics.median(timed_region(rm, [g]) for _ in range(3)) * 1e3 / (Br * Tr)
                rollout[f"{nn}_{mode}"] = {"num_envs": nn, "batches": Br, "us_per_step": us}
            assert max(e.status() for e in re_) == 0
            for e in re_:
                e.close()
            del re_, outs, racts
            torch.cuda.empty_cache()

    # ---- end to end through the public API with HOST buffers (numpy in, numpy out)
    e2e_steps = args.e2e_steps
    # env batches in flight: three on one or two GPUs (16 host cores per GPU: the decode of one batch runs on the library's host pool while
    # the caller enqueues the others); ONE from four GPUs up - there a rank has 4 cores, the eight ranks' observation mirrors (8 x 19.6 MB)
    # no longer fit the host's caches, and a second batch in flight per rank only doubles that working set (measured on 8 GPUs:
    # 8.4e8 env-steps/s with one batch in flight against 5.2-6.2e8 with two, tools/dev/e2e_multi8.sh)
    # measured on this pool's boxes (16 / 24 / 32 / 32 vCPUs for 1 / 2 / 4 / 8 GPUs): the ranks share the host's cores and memory system,
    # so the best number of batches in flight per rank falls as ranks are added (N=1: 0.68 / 0.90 / 0.87e9 with 2 / 3 / 4; N=2: 0.64 / 1.05 /
    # 0.92e9 with 1 / 2 / 3; N=4: 1.01 / 0.84 / 0.62e9 with 1 / 2 / 3)
    EB = min(B, args.e2e_batches if args.e2e_batches > 0 else (3 if world <= 1 else (2 if world == 2 else 1)))
    host_rings = [rings[b].cpu().numpy() for b in range(EB)]
    cells = envs[0].width * envs[0].height

    def e2e_run(transport, pipelined):
        for b in range(EB):
            envs[b].set_host_transport(transport)
        for i in range(2 * EB + 1):   # first call per env allocates its page-locked buffers: keep that out of the timing
            envs[i % EB].step(host_rings[i % EB][i % RING])
        barrier()
        chk = 0.0
        t0 = time.perf_counter()
        if not pipelined:
            for i in range(e2e_steps):
                obs, rew, term, trunc, _ = envs[i % EB].step(host_rings[i % EB][(i // EB) % RING])
                chk += float(rew[0, 0]) + float(obs[0, 1, 8, 0])
        else:
            # the same calls split in their two halves (VectorEnv.step_async / step_wait) with EB env batches in flight: every
            # step still carries its own H2D of actions and D2H of results, but the copies and the host-side decode of one
            # batch overlap the step of the other
            for b in range(EB):
                envs[b].step_async(host_rings[b][0])
            for i in range(e2e_steps):
                b = i % EB
                obs, rew, term, trunc, _ = envs[b].step_wait()
                chk += float(rew[0, 0]) + float(obs[0, 1, 8, 0])
                if i + EB < e2e_steps:
                    envs[b].step_async(host_rings[b][((i + EB) // EB) % RING])
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, chk

    rec_bytes = 16 + n * 16       # delta records: 16 B / env (2 agents, 10x10) + the 16-byte header carrying the autoreset count
    e2e = {}
    for transport, d2h in (("delta", rec_bytes + n * (cells + 4) // 50), ("packed", n * (cells + 16 + 2)), ("full", n * (3 * cells + 16 + 2))):
        secs = [e2e_run(transport, True)[0] for _ in range(args.e2e_repeats)]
        e2e[transport] = {"s": statistics.median(secs), "d2h": d2h}
    block_s = statistics.median(e2e_run("delta", False)[0] for _ in range(args.e2e_repeats))
    for b in range(EB):
        envs[b].set_host_transport("delta")

    # for scale: host actions in, host rewards / flags out, observations left on the device (where a GPU policy reads them)
    pin_act = [torch.as_tensor(host_rings[b][0]).pin_memory() for b in range(EB)]
    pin_rew = torch.empty((n, 2), dtype=torch.float64, pin_memory=True)
    pin_flags = torch.empty((2, n), dtype=torch.bool, pin_memory=True)

    def dev_obs_step(b):
        o, r, te, tr, _ = envs[b].step(pin_act[b].to(dev, non_blocking=True))
        pin_rew.copy_(r, non_blocking=True); pin_flags[0].copy_(te, non_blocking=True); pin_flags[1].copy_(tr, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return float(pin_rew[0, 0])
    for i in range(4):
        dev_obs_step(i % EB)
    dsteps = e2e_steps * 2
    t0 = time.perf_counter()
    for i in range(dsteps):
        dev_obs_step(i % EB)
    e2e_devobs_s = (time.perf_counter() - t0) / dsteps

    # for scale: a bare device-to-host copy of one FULL step result (318 B / env) into page-locked memory on this box
    dsrc = torch.empty(n * (3 * cells + 16 + 2), dtype=torch.uint8, device=dev)
    hdst = torch.empty(n * (3 * cells + 16 + 2), dtype=torch.uint8, pin_memory=True)
    for _ in range(3):
        hdst.copy_(dsrc, non_blocking=True)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    for _ in range(20):
        hdst.copy_(dsrc, non_blocking=True)
    torch.cuda.synchronize(dev)
    bare_d2h_gbps = 20 * dsrc.numel() / (time.perf_counter() - t1) / 1e9
    del dsrc, hdst
    for e in envs:
        e.close()
    del envs, rings
    torch.cuda.empty_cache()

    families = {} if args.skip_families else bench_families(args, dev, rank, world, timed_region_factory=(barrier, sampler))

    fam_keys = sorted(families)
    roll_keys = sorted(k for k in (rollout or {}) if isinstance(rollout[k], dict))
    vec = [rollout[k]["us_per_step"] for k in roll_keys] + [ms, single_us, e2e["delta"]["s"] * 1e3, e2e["packed"]["s"] * 1e3, e2e["full"]["s"] * 1e3, block_s * 1e3, e2e_devobs_s * 1e3,
           (small or {}).get("avg_launch_us", 0.0)] + [families[k]["us_per_step"] for k in fam_keys]
    t = torch.tensor(vec, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vals = [float(x) for x in t]
    for k in roll_keys:
        rollout[k]["us_per_step"] = vals.pop(0)
    ms_max, single_us_max, e2e_ms, e2e_packed_ms, e2e_full_ms, e2e_block_ms, e2e_devobs_ms, small_us = vals[:8]
    for k, v in zip(fam_keys, vals[8:]):
        families[k]["us_per_step"] = v
    sampler.stop_flag = True

    if rank == 0:
        peak, peak_src = measured_peak()
        value = K * n * world / (ms_max * 1e-3)
        launch_s = ms_max * 1e-3 / K
        achieved = ALGO_BYTES_PER_ENV_STEP * n / launch_s / 1e9
        d2h_delta = e2e["delta"]["d2h"]
        for k in fam_keys:
            f = families[k]
            f["ms_per_step"] = f["us_per_step"] / 1e3
            f["value"] = f["num_envs_per_gpu"] * world / (f["us_per_step"] * 1e-6)
            f["unit"] = UNIT
            ach = f["algorithmic_bytes_per_env_step"] * f["num_envs_per_gpu"] / (f["us_per_step"] * 1e-6) / 1e9
            f["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "algorithmic_bytes": f["algorithmic_bytes_per_env_step"] * f["num_envs_per_gpu"]}
        for k in roll_keys:
            rr = rollout[k]
            rr["value"] = rr["num_envs"] * world / (rr["us_per_step"] * 1e-6)
            rr["frac_592B"] = ALGO_BYTES_PER_ENV_STEP * rr["num_envs"] / rr["us_per_step"] / 1e3 / peak
            rr["frac_moved"] = rollout["bytes_moved_per_env_step"] * rr["num_envs"] / rr["us_per_step"] / 1e3 / peak
        if small:
            small.update(avg_launch_us=small_us, value=small["num_envs"] * world / small_us * 1e6,
                         frac=ALGO_BYTES_PER_ENV_STEP * small["num_envs"] / small_us / 1e3 / peak)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "agent_steps_per_sec": value * 2,
            "config": workload_config(args),
            "repeats": R, "repeat_ms": all_ms, "pickups_per_env_step": pickups,
            "launch": f"CUDA graphs of up to {CHUNK} launches of the fused step+encode kernel (one launch = one env batch of {n} envs), "
                      f"the {B} independent batches forked over {S} streams; value = median of {R} timed regions of exactly {K} launches, each timed "
                      "with CUDA events on the launching stream right behind an untimed 160 MB fill (L2 flushed; the host's graph-launch "
                      "latency stays outside the device-timed interval)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "collect_rollout_kernel (warp-tile kernel, T = 1: mg_step)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n,
                         "avg_launch_us": launch_s * 1e6,
                         "traffic_frac": (ncu_traffic() / launch_s / 1e9 / peak) if (ncu_traffic() and n == 65536) else None,
                         "note": "achieved = SURVEY 8(d)'s 592 algorithmic B/env-step x envs per launch / average launch time over the timed "
                                 "region (launches of independent env batches overlap across streams); the packed layout moves fewer bytes "
                                 "than that count (traffic = DRAM bytes per launch from the committed ncu capture). single_stream = the "
                                 "same launches serialised on one stream"},
            "single_stream": {"avg_launch_us": single_us_max, "value": n * world / single_us_max * 1e6,
                              "achieved": ALGO_BYTES_PER_ENV_STEP * n / single_us_max / 1e3,
                              "frac": ALGO_BYTES_PER_ENV_STEP * n / single_us_max / 1e3 / peak, "launches": n1,
                              "note": "same launch sequence on ONE stream (launches serialised, programmatic dependent launch)"},
            "small_batch": small,
            "rollout": rollout,
            "e2e": {"value": e2e_steps * n * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * 2, "d2h_bytes_per_step": d2h_delta,
                    "steps": e2e_steps, "repeats": args.e2e_repeats, "transport": "delta",
                    "api": f"CollectVecEnv.step_async(numpy) / step_wait() -> mg_step_host_async / _wait, {EB} env batches in flight, fresh actions per step; "
                           "the observation crosses PCIe as 16-byte per-env records of the cells the step changed (+ the packed rows of the "
                           "envs that autoreset: every 50th step here, averaged into d2h_bytes_per_step) and is patched into the page-locked "
                           f"(N, W, H, 3) uint8 array the call returns by {host_threads() // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1'))) or 1} host threads "
                           "(the library's pool: the decode of a batch starts when its copy has landed and overlaps the caller's enqueue of the "
                           "other batches; the steady-state step goes out as one CUDA-graph launch)",
                    "d2h_GBps_per_gpu": d2h_delta / (e2e_ms * 1e-3 / e2e_steps) / 1e9,
                    "bare_d2h_copy_GBps": bare_d2h_gbps,
                    "packed": {"value": e2e_steps * n * world / (e2e_packed_ms * 1e-3), "d2h_bytes_per_step": e2e["packed"]["d2h"],
                               "api": "same calls, MG_TRANSPORT_PACKED: the 1-byte-per-cell grid plane + rewards + flags, expanded on the host"},
                    "full": {"value": e2e_steps * n * world / (e2e_full_ms * 1e-3), "d2h_bytes_per_step": e2e["full"]["d2h"],
                             "api": "same calls, MG_TRANSPORT_FULL: the expanded observation (round 1's path)",
                             "d2h_GBps_per_gpu": e2e["full"]["d2h"] / (e2e_full_ms * 1e-3 / e2e_steps) / 1e9},
                    "blocking": {"value": e2e_steps * n * world / (e2e_block_ms * 1e-3),
                                 "api": "CollectVecEnv.step(numpy) -> mg_step_host (delta transport), one call at a time"},
                    "obs_on_device": {"value": n * world / (e2e_devobs_ms * 1e-3), "h2d_bytes_per_step": n * 2, "d2h_bytes_per_step": n * (16 + 2),
                                      "api": "CollectVecEnv.step(cuda tensor): pinned actions H2D, rewards + flags D2H, obs stay in HBM for a GPU policy"}},
            "families": families,
            "gpu_launches": K,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            nthreads = host_threads()
            cpu_steps = args.cpu_steps
            v, dt, pk = time_oracle(n, cpu_steps, 2, nthreads, ring=RING)
            if dt < 5.0:   # aim for ~10-30 s of CPU work
                cpu_steps = int(cpu_steps * 12.0 / max(dt, 1e-3))
                v, dt, pk = time_oracle(n, cpu_steps, 1, nthreads, ring=RING)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "pickups_per_env_step": pk,
                                    "sample": f"{cpu_steps} steps x {n} envs ({dt:.1f} s), same config and fresh actions per step, "
                                              "oracle/mg_oracle.c with OpenMP over envs",
                                    "python_reference": python_reference(args.python_seconds)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def bench_families(args, dev, rank, world, timed_region_factory):
    """BASELINE configs 3 / 4 / 5 on this GPU, same protocol (CUDA graph over env batches larger than L2, fresh actions per step
    from a ring, CUDA events, median of 3 regions): CtF 2v2 on the reference's test board, Maze 64x64 fused step + V = 7 partial
    views, Wildfire 64x64 x 16 agents.  Maps come from the committed golden fixtures (nothing under /root/reference is read)."""
    import numpy as np
    import torch

    import gym_multigrid_b200 as mg
    barrier, sampler = timed_region_factory
    golden = os.path.join(ROOT, "tests", "golden")
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    out = {}

    def run(name, envs, n_act, act_shape, steps, bytes_per_env, n, note, warm=0):
        ring = 8
        acts = [torch.randint(0, n_act, (ring,) + act_shape, generator=gen, device=dev, dtype=torch.int8) for _ in envs]
        Bf = len(envs)
        for e in envs:
            e.reset()
        torch.cuda.synchronize(dev)     # the resets ran on the default stream; `main` is a non-blocking stream
        main = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(main):
            for i in range(max(warm, 3) * Bf):
                envs[i % Bf].step(acts[i % Bf][(i // Bf) % ring])
            main.synchronize()
            count = Bf * ring
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                for gi in range(count):
                    envs[gi % Bf].step(acts[gi % Bf][(gi // Bf) % ring])
            g.replay()
            main.synchronize()
        reps = max(1, -(-steps // count))
        samples = []
        for _ in range(3):
            barrier()
            with torch.cuda.stream(main):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                sampler.sample()
                e0.record(main)
                for _ in range(reps):
                    g.replay()
                e1.record(main)
                main.synchronize()
                sampler.sample()
            samples.append(e0.elapsed_time(e1) * 1e3 / (reps * count))
        st = max((e.status() if hasattr(e, "status") else 0) for e in envs)
        assert st == 0, f"{name}: device status word {st}"
        out[name] = {"workload": note, "num_envs_per_gpu": n, "env_batches": Bf, "steps_timed": reps * count, "us_per_step": statistics.median(samples),
                     "algorithmic_bytes_per_env_step": bytes_per_env, "actions": f"fresh per step (ring of {ring})"}
        for e in envs:
            e.close()
        torch.cuda.empty_cache()

    with np.load(os.path.join(golden, "ctf_2v2.npz")) as z:
        board = z["field_map"]
    with np.load(os.path.join(golden, "maze_gen64.npz")) as z:
        maze = z["field_map"]
    base = (world + rank) << 32          # global env ids disjoint from the Collect batches and between ranks
    n = args.family_envs
    nb, nr = 2, 2
    run("ctf_2v2", [mg.make_ctf_vec(n, board, num_blue_agents=nb, num_red_agents=nr, seed=args.seed, env_id_base=base + b * n, device=dev) for b in range(2)],
        5, (n, nb), 160, 100 + 2 * (4 * (nb + nr) + 16) + nb + 10, n,
        "BASELINE config 3: CtFMvNEnv 2v2 on tests/assets/board.txt (10x10), RwPolicy reds drawn on the device, u8 map observation, max_steps 100 + autoreset")
    mz = [mg.make_maze_vec(n, maze, seed=args.seed, env_id_base=base + (2 + b) * n, device=dev) for b in range(2)]
    for e in mz:
        e.set_partial_obs(7)
    run("maze64_partial7", mz, 5, (n,), 160, 147 + 2 * (4 + 16) + 1 + 10, n,
        "BASELINE config 4: MazeSingleAgentEnv on a generated 64x64 map, fused step + V=7 partial-view observation (one launch), autoreset; "
        "Maze + partial view is a composition the reference does not ship: cells outside the map show the filler (3, 7, 1), an extension (unpinned for that one code)")
    nw, A, size = args.wildfire_envs, 16, 64
    cells = size * size
    run("wildfire64_a16", [mg.make_wildfire_vec(nw, size=size, num_agents=A, seed=args.seed, env_id_base=base + (4 + b) * n, device=dev) for b in range(2)],
        5, (nw, A), 32, A + 2 * (cells + 4 * A + 16) + 3 * cells + 8 * A + 2, nw,
        "BASELINE config 5: Wildfire 64x64, 16 agents (extension: no reference code; own spec + own oracle, parity unpinned)", warm=30)
    return out


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else a library prints there (NCCL's version
    banner, torchrun notices) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)          # stray prints of native libraries on fd 1 -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--batches", type=int, default=16)
    ap.add_argument("--action-ring", type=int, default=64, help="independently drawn action tensors per env batch, one per step in turn")
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of exactly --steps launches; the median is reported")
    ap.add_argument("--streams", type=int, default=4, help="streams the independent env batches are forked over inside the graph")
    ap.add_argument("--e2e-steps", type=int, default=384)
    ap.add_argument("--e2e-batches", type=int, default=0, help="env batches in flight in the e2e loop (0 = 3 on one GPU, 2 on two, 1 from four up)")
    ap.add_argument("--e2e-repeats", type=int, default=3)
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--python-seconds", type=float, default=3.0, help="seconds per leg of the Python-reference baseline (when a copy is on the box)")
    ap.add_argument("--family-envs", type=int, default=1 << 20, help="CtF / Maze envs per GPU in the `families` lines")
    ap.add_argument("--wildfire-envs", type=int, default=131072)
    ap.add_argument("--skip-families", action="store_true")
    ap.add_argument("--skip-small", action="store_true")
    ap.add_argument("--skip-rollout", action="store_true")
    ap.add_argument("--rollout-steps", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--budget-s", type=float, default=60.0, help="wall budget of the --impl reference arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
