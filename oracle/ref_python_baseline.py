"""Throughput of the UNMODIFIED Python reference on host cores, reproducible in the build container.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: needs the reference package - /root/reference (build container) or the copy that
oracle/install_reference.py pip-installs into baseline/_ref (travels to the GPU box; set MG_REFERENCE_ROOT to it) - and imports
it under oracle/refshim exactly as the golden recorder does.  Nothing in the product or the tests imports this file; bench.py's
CPU-baseline legs call `measure()` to quote the Python reference beside the C port (`cpu_baseline.python_reference`).

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


def _other_families(seconds):
    """Single-process step rates of the reference's CtF (2v2, RwPolicy reds, each observation option) and Maze (13x13 test board
    and a 64x64 map) envs, and the cost of one `render()` call (tile cache warm) - context for BASELINE configs 3 and 4."""
    import tempfile

    import ref_harness as rh
    rh.import_reference()
    from gym_multigrid.envs.ctf import CtFMvNEnv
    from gym_multigrid.envs.maze import MazeSingleAgentEnv
    from gym_multigrid.policy.ctf.heuristic import RwPolicy
    assets = os.path.join(rh.REFERENCE_ROOT, "tests", "assets")
    if not os.path.isdir(assets):
        assets = os.path.join(rh.REFERENCE_ROOT, "assets")     # the copy installed by oracle/install_reference.py
    out = {}

    def rate(env, sample, seconds):
        env.reset(seed=0)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            _, _, term, trunc, _ = env.step(sample())
            n += 1
            if term or trunc:
                env.reset()
        return n / (time.perf_counter() - t0)

    rng = np.random.default_rng(0)
    for opt in ("map", "flattened", "positional"):
        env = CtFMvNEnv(os.path.join(assets, "board.txt"), num_blue_agents=2, num_red_agents=2, enemy_policies=RwPolicy(), observation_option=opt)
        out[f"ctf_2v2_{opt}_env_steps_per_s"] = rate(env, lambda: [int(a) for a in rng.integers(0, 5, size=2)], seconds)
    env.reset(seed=0); env.render()
    t0 = time.perf_counter()
    for _ in range(20):
        env.step([int(a) for a in rng.integers(0, 5, size=2)]); env.render()
    out["ctf_2v2_step_plus_render_per_s"] = 20 / (time.perf_counter() - t0)
    np.random.seed(0)
    env = MazeSingleAgentEnv(os.path.join(assets, "board_maze.txt"), observation_option="map")
    out["maze_13x13_env_steps_per_s"] = rate(env, lambda: int(rng.integers(0, 5)), seconds)
    m = (np.random.default_rng(0).random((64, 64)) < 0.2).astype(np.int64) * 3
    m[5, 7] = 2
    tmp = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    np.savetxt(tmp.name, m.T, fmt="%d")
    env = MazeSingleAgentEnv(tmp.name, observation_option="map")
    out["maze_64x64_env_steps_per_s"] = rate(env, lambda: int(rng.integers(0, 5)), seconds)
    os.unlink(tmp.name)
    return out


def measure(seconds=3.0, workers=None, families=True):
    """The numbers as a dict (see the module docstring); `workers` defaults to the cores of this process's affinity mask."""
    if workers is None:
        try:
            workers = len(os.sched_getaffinity(0))
        except AttributeError:
            workers = os.cpu_count() or 1
    single = _loop(seconds, 0)
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_indep, args=(seconds, 100 + i, q)) for i in range(workers)]
    [p.start() for p in procs]
    indep = sum(q.get() for _ in procs)
    [p.join() for p in procs]
    pipes = [ctx.Pipe() for _ in range(workers)]
    ws = [ctx.Process(target=_worker, args=(c, 200 + i), daemon=True) for i, (_, c) in enumerate(pipes)]
    [w.start() for w in ws]
    rng = np.random.default_rng(1)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        acts = rng.integers(0, 4, size=(workers, 2))
        for (p, _), a in zip(pipes, acts):
            p.send([int(a[0]), int(a[1])])
        for p, _ in pipes:
            p.recv()
        steps += workers
    lockstep = steps / (time.perf_counter() - t0)
    for p, _ in pipes:
        p.send(None)
    [w.join(timeout=5) for w in ws]
    out = {"env": ENV_ID, "impl": "unmodified Python reference under oracle/refshim", "cores": workers,
           "single_process_env_steps_per_s": single, "independent_processes_env_steps_per_s": indep,
           "lockstep_pipes_env_steps_per_s": lockstep, "seconds_each": seconds}
    if families:
        out["other_families_single_process"] = _other_families(seconds)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--no-families", action="store_true")
    args = ap.parse_args()
    print(json.dumps(measure(args.seconds, args.workers, not args.no_families)))


if __name__ == "__main__":
    main()
