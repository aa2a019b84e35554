#!/usr/bin/env python
"""Benchmark of the hot path: batched `CollectGameEnv.step` + `Grid.encode` (BASELINE.json configs[1]:
multigrid-collect-respawn-clustered-v0, 2 agents, 65 536 envs per launch per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one fused step+encode launch over one batch of 65 536 envs.  To defeat the 126 MB L2
the timed loop rotates over `--batches` independent env batches (each with its own state, obs and
action tensors; 16 x 39 MB = 630 MB working set), so every launch streams its state from HBM.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions used here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "multigrid-collect-respawn-clustered-v0"
ALGO_BYTES_PER_ENV_STEP = 592   # actions 2 + state 2x136 + obs 300 + f64 rewards 16 + flags 2
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


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
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.active, self.sm, self.reasons, self.max_mhz = index, False, False, [], set(), None
        self.nv = self.h = None
        try:    # NVML is initialised HERE (caller's thread, before the timed region): the thread only polls
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        nv, h = self.nv, self.h
        if nv is None:
            return
        try:
            names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                if self.active:     # samples are kept only while the timed region runs
                    self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = get_reasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.001)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_oracle(num_envs, steps, warmup, nthreads, seed=0):
    """CPU restatement of the reference's path (oracle/mg_oracle.c, OpenMP over envs), same config."""
    import numpy as np
    o, r = oracle_env(num_envs, nthreads, seed)
    acts = np.random.default_rng(seed).integers(0, 4, size=(num_envs, 2)).astype(np.int8)
    for _ in range(warmup):
        o.step(acts, r, autoreset=True, reuse_buffers=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(acts, r, autoreset=True, reuse_buffers=True)
    dt = time.perf_counter() - t0
    return num_envs * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = host_threads()
    # bounded sample: keep EXACTLY --steps steps, size the envs per step so the run ends in ~1 minute
    probe, _ = time_oracle(8192, 20, 2, nthreads)
    n = int(min(args.num_envs, max(512, probe * args.budget_s / max(1, args.steps + args.warmup))))
    n = max(64, n // 64 * 64)
    value, dt = time_oracle(n, args.steps, args.warmup, nthreads)
    sample = (f"{args.steps} steps x {n} envs per step ({dt:.1f} s; GPU arm steps {args.num_envs} envs per launch), "
              "autoreset, Philox RNG, OpenMP over envs")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "agent_steps_per_sec": value * 2,
        "config": {"workload": f"{ENV_ID}, 2 agents, {n} envs per step, uniform random actions, TimeLimit 50 + same-step autoreset",
                   "num_envs": n, "note": "CPU restatement (C, oracle/mg_oracle.c) of the reference's Python path on the host cores; "
                                          "the Python reference itself cannot travel to this box (it measured ~7.2e3 env-steps/s/core, BASELINE.md)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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

    n, B = args.num_envs, args.batches
    K, Wm = args.steps, max(3, args.warmup)
    envs, acts = [], []
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    for b in range(B):
        e = mg.make_vec(ENV_ID, n, device=dev, seed=args.seed, autoreset=True, env_id_base=(rank * B + b) * n)
        e.reset()
        envs.append(e)
        acts.append(torch.randint(0, 4, (n, 2), generator=gen, device=dev, dtype=torch.int8))
    torch.cuda.synchronize(dev)

    # ---- device-resident throughput: CUDA graph of B launches (one per batch), replayed.  The B env batches are
    #      independent, so the graph forks them over `--streams` streams: tiles of one launch load while tiles of
    #      another compute / drain (a single stream serialises whole launches, which leaves HBM idle during every
    #      launch's ramp-up and store drain; that figure is reported beside it as `single_stream`).
    def capture(n_streams, main=None, count=None):
        """Graph of `count` (default B) launches, batch b on stream b % n_streams."""
        count = B if count is None else count
        fresh = main is None
        main = main or torch.cuda.Stream(device=dev)
        side = [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
        with torch.cuda.stream(main):
            if fresh:
                for i in range(Wm):
                    envs[i % B].step(acts[i % B])
            main.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                for s_ in side:
                    s_.wait_stream(main)
                for i in range(count):
                    b = i % B                      # launches i and i + B step the same env batch: same stream, in order
                    st = main if b % n_streams == 0 else side[b % n_streams - 1]
                    with torch.cuda.stream(st):
                        envs[b].step(acts[b])
                for s_ in side:
                    main.wait_stream(s_)
        return main, g

    def timed(main, g, reps, sample_clocks, big=None, tail=None):
        """Warm up with the B-launch graph `g`, then time `reps` replays of `big` (default `g`) + one replay of `tail`."""
        big = big or g
        smp = ClockSampler(local_rank) if sample_clocks else None
        if smp:
            smp.start()         # polling thread up and NVML initialised before the warm-up; it records only while `active`
        with torch.cuda.stream(main):
            for _ in range(max(200, Wm // B)):     # untimed: the W warm-up steps and ~20 ms more, so a short timed region (small --steps) runs at settled clocks
                g.replay()
            for x in (big, tail):
                if x is not None and x is not g:
                    x.replay()                     # a graph's first launch uploads it: keep that out of the timed region
            main.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if smp:
                smp.active = True
            ev0.record(main)
            for _ in range(reps):
                big.replay()
            if tail is not None:
                tail.replay()
            ev1.record(main)
            main.synchronize()
            torch.cuda.synchronize(dev)
            if smp:
                smp.active = False
                smp.stop_flag = True
                smp.join(timeout=2)
        return ev0.elapsed_time(ev1), smp

    # EXACTLY K timed steps.  The timed graphs hold up to CHUNK launches each (launch i steps env batch i % B on stream
    # (i % B) % S, so launches of one batch stay in order on one stream and the S streams only join at the end of a graph):
    # K // CHUNK replays of the CHUNK-launch graph + one graph of the K % CHUNK remaining launches.
    CHUNK = 4096 - 4096 % B
    K_eff = max(1, K)
    reps, rem = divmod(K_eff, CHUNK)
    S = max(1, min(args.streams, B))
    main1, graph1 = capture(1)
    n1 = max(B, min(K_eff, 1024))
    ms_single, _ = timed(main1, graph1, 1, False, big=capture(1, main=main1, count=n1)[1])
    single_us = ms_single * 1e3 / n1
    mainS, graphS = (main1, graph1) if S == 1 else capture(S)
    bigS = capture(S, main=mainS, count=CHUNK)[1] if reps else None
    tailS = capture(S, main=mainS, count=rem)[1] if rem else None
    ms, sampler = timed(mainS, graphS, reps, True, big=bigS, tail=tailS)
    status = max(e.status() for e in envs)
    assert status == 0, f"device status word {status}"

    # ---- end to end through the public API with HOST buffers (numpy in, numpy out)
    e2e_steps = args.e2e_steps
    host_act = [a.cpu().numpy() for a in acts]
    EB = min(B, 2)   # PCIe-bound: L2 residency is irrelevant here, two batches keep page-locked memory small
    for i in range(2 * EB + 1):   # first call per env allocates its page-locked buffers: keep that out of the timing
        envs[i % EB].step(host_act[i % EB])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(e2e_steps):
        obs, rew, term, trunc, _ = envs[i % EB].step(host_act[i % EB])
        checksum += float(rew[0, 0]) + float(obs[0, 1, 8, 0])
    torch.cuda.synchronize(dev)
    e2e_block_s = time.perf_counter() - t0

    # the same calls split in their two halves (VectorEnv.step_async / step_wait) with EB env batches in flight: every step
    # still carries its own H2D of actions and D2H of obs / rewards / flags, but the result copy of one batch overlaps the
    # step of the other, so the PCIe link never idles.  This is the headline e2e; the blocking figure is reported beside it.
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for b in range(EB):
        envs[b].step_async(host_act[b])
    for i in range(e2e_steps):
        b = i % EB
        obs, rew, term, trunc, _ = envs[b].step_wait()
        checksum += float(rew[0, 0]) + float(obs[0, 1, 8, 0])
        if i + EB < e2e_steps:
            envs[b].step_async(host_act[b])
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # for scale: host actions in, host rewards / flags out, observations left on the device (where a GPU policy reads them)
    pin_act = [torch.as_tensor(a).pin_memory() for a in host_act[:EB]]
    pin_rew = torch.empty((n, 2), dtype=torch.float64, pin_memory=True)
    pin_flags = torch.empty((2, n), dtype=torch.bool, pin_memory=True)
    def dev_obs_step(b):
        o, r, te, tr, _ = envs[b].step(pin_act[b].to(dev, non_blocking=True))
        pin_rew.copy_(r, non_blocking=True); pin_flags[0].copy_(te, non_blocking=True); pin_flags[1].copy_(tr, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return float(pin_rew[0, 0])
    for i in range(4):
        dev_obs_step(i % EB)
    dsteps = e2e_steps * 4
    t0 = time.perf_counter()
    for i in range(dsteps):
        checksum += dev_obs_step(i % EB)
    e2e_devobs_s = (time.perf_counter() - t0) / dsteps

    # for scale: a bare device-to-host copy of one step's result bytes into page-locked memory on this box
    dsrc = torch.empty(n * (300 + 16 + 2), dtype=torch.uint8, device=dev)
    hdst = torch.empty(n * (300 + 16 + 2), dtype=torch.uint8, pin_memory=True)
    for _ in range(3):
        hdst.copy_(dsrc, non_blocking=True)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    for _ in range(20):
        hdst.copy_(dsrc, non_blocking=True)
    torch.cuda.synchronize(dev)
    bare_d2h_gbps = 20 * dsrc.numel() / (time.perf_counter() - t1) / 1e9
    del dsrc, hdst

    t = torch.tensor([ms, e2e_s * 1e3, e2e_block_s * 1e3, e2e_devobs_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e_block_ms_max, e2e_devobs_ms_max = (float(x) for x in t)

    if rank == 0:
        value = K_eff * n * world / (ms_max * 1e-3)
        e2e_value = e2e_steps * n * world / (e2e_ms_max * 1e-3)
        peak, peak_src = measured_peak()
        launch_s = ms * 1e-3 / K_eff
        achieved = ALGO_BYTES_PER_ENV_STEP * n / launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_eff, "warmup": Wm,
            "ms_per_step": ms_max / K_eff, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "agent_steps_per_sec": value * 2,
            "config": {"workload": f"{ENV_ID}, 2 agents, {n} envs per launch per GPU, uniform random actions, TimeLimit 50 + same-step autoreset, Philox RNG",
                       "num_envs_per_gpu_per_launch": n, "env_batches_per_gpu": B,
                       "l2": f"inputs larger than L2: timed loop rotates over {B} independent env batches "
                             f"({B * n * (ALGO_BYTES_PER_ENV_STEP + 8) / 1e6:.0f} MB working set > 126 MB L2)",
                       "launch": f"CUDA graphs of up to {CHUNK} launches of the fused step+encode kernel (one launch = one env batch), "
                                 f"the {B} independent batches forked over {S} streams",
                       "streams": S},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "collect_step_kernel",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n,
                         "avg_launch_us": launch_s * 1e6,
                         "traffic_frac": (ncu_traffic() / launch_s / 1e9 / peak) if (ncu_traffic() and n == 65536) else None,
                         "note": "achieved = SURVEY 8(d)'s 592 algorithmic B/env-step x envs per launch / average launch time; the packed "
                                 "layout moves fewer bytes than that count (traffic = DRAM bytes per launch from the committed ncu capture), "
                                 "so frac can exceed 1 while traffic_frac = traffic / launch time / peak stays below it"},
            "single_stream": {"avg_launch_us": single_us, "value": n / single_us * 1e6,
                              "achieved": ALGO_BYTES_PER_ENV_STEP * n / single_us / 1e3, "frac": ALGO_BYTES_PER_ENV_STEP * n / single_us / 1e3 / peak,
                              "note": "same graph on ONE stream (launches serialised by programmatic dependent launch)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 2, "d2h_bytes_per_step": n * (300 + 16 + 2),
                    "steps": e2e_steps,
                    "api": f"CollectVecEnv.step_async(numpy) / step_wait() -> mg_step_host_async / _wait (pinned host buffers), {EB} env batches in flight",
                    "d2h_GBps_per_gpu": n * (300 + 16 + 2) / (e2e_ms_max * 1e-3 / e2e_steps) / 1e9,
                    "bare_d2h_copy_GBps": bare_d2h_gbps,
                    "blocking": {"value": e2e_steps * n * world / (e2e_block_ms_max * 1e-3),
                                 "api": "CollectVecEnv.step(numpy) -> mg_step_host, one call at a time",
                                 "d2h_GBps_per_gpu": n * (300 + 16 + 2) / (e2e_block_ms_max * 1e-3 / e2e_steps) / 1e9},
                    "obs_on_device": {"value": n * world / (e2e_devobs_ms_max * 1e-3), "h2d_bytes_per_step": n * 2, "d2h_bytes_per_step": n * (16 + 2),
                                      "api": "CollectVecEnv.step(cuda tensor): pinned actions H2D, rewards + flags D2H, obs stay in HBM for a GPU policy"},
                    "note": "bound by the device-to-host copy of the observations (300 B/env over PCIe), not by the kernel; "
                            "bare_d2h_copy_GBps = the same bytes copied by torch alone on this box"},
            "gpu_launches": K_eff,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            nthreads = host_threads()
            cpu_steps = args.cpu_steps
            v, dt = time_oracle(n, cpu_steps, 2, nthreads)
            if dt < 5.0:   # aim for ~10-30 s of CPU work
                cpu_steps = int(cpu_steps * 12.0 / max(dt, 1e-3))
                v, dt = time_oracle(n, cpu_steps, 1, nthreads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port",
                                    "sample": f"{cpu_steps} steps x {n} envs ({dt:.1f} s), same config, oracle/mg_oracle.c with OpenMP over envs"}
        emit(line)
    for e in envs:
        e.close()
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--streams", type=int, default=4, help="streams the independent env batches are forked over inside the graph")
    ap.add_argument("--e2e-steps", type=int, default=64)
    ap.add_argument("--cpu-steps", type=int, default=40)
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
