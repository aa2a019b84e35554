"""Throughput of the UNMODIFIED Python reference on host cores, reproducible in the build container.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: needs /root/reference (absent on the GPU box), imports the reference under
oracle/refshim exactly as the golden recorder does.  Nothing in the product, the tests or bench.py imports this file.

    python oracle/ref_python_baseline.py [--seconds 3] [--workers N]

Prints one JSON line: single-process step-only rate, N independent processes (no IPC: the upper bound of what host cores
can do), and N worker processes stepped in lockstep over pipes (the gymnasium AsyncVectorEnv pattern: one env per process)."""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
ENV_ID = "multigrid-collect-respawn-clustered-v0"


def _make():
    import ref_harness as rh
    rh.import_reference()
    env, time_limit = rh.make_collect(ENV_ID)
    return env, time_limit


def _loop(seconds, seed):
    """50-step episodes (the registration's TimeLimit) incl. reset(), uniform random actions; returns env-steps/s."""
    env, tl = _make()
    random.seed(seed); np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env.reset()
    steps, t0, k = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        _, _, term, trunc, _ = env.step([int(a) for a in rng.integers(0, 4, size=2)])
        steps += 1; k += 1
        if term or trunc or k >= tl:
            env.reset(); k = 0
    return steps / (time.perf_counter() - t0)


def _indep(seconds, seed, q):
    q.put(_loop(seconds, seed))


def _worker(conn, seed):
    env, tl = _make()
    random.seed(seed); np.random.seed(seed)
    env.reset()
    k = 0
    while True:
        msg = conn.recv()
        if msg is None:
            return
        obs, rew, term, trunc, _ = env.step(msg)
        k += 1
        if term or trunc or k >= tl:
            obs, _ = env.reset(); k = 0
        conn.send((obs, rew, term, trunc))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    single = _loop(args.seconds, 0)
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_indep, args=(args.seconds, 100 + i, q)) for i in range(args.workers)]
    [p.start() for p in procs]
    indep = sum(q.get() for _ in procs)
    [p.join() for p in procs]
    pipes = [ctx.Pipe() for _ in range(args.workers)]
    workers = [ctx.Process(target=_worker, args=(c, 200 + i), daemon=True) for i, (_, c) in enumerate(pipes)]
    [w.start() for w in workers]
    rng = np.random.default_rng(1)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < args.seconds:
        acts = rng.integers(0, 4, size=(args.workers, 2))
        for (p, _), a in zip(pipes, acts):
            p.send([int(a[0]), int(a[1])])
        for p, _ in pipes:
            p.recv()
        steps += args.workers
    lockstep = steps / (time.perf_counter() - t0)
    for p, _ in pipes:
        p.send(None)
    print(json.dumps({"env": ENV_ID, "impl": "unmodified Python reference under oracle/refshim", "cores": args.workers,
                      "single_process_env_steps_per_s": single, "independent_processes_env_steps_per_s": indep,
                      "lockstep_pipes_env_steps_per_s": lockstep, "seconds_each": args.seconds}))


if __name__ == "__main__":
    main()
