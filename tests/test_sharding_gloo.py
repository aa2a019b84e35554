"""N>1 host path on CPU: world_size-2 gloo processes.  Each rank steps ITS shard (through the
oracle, since there is no GPU here) with Philox streams keyed by global env id; the concatenation
must equal the unsharded run, and the stat reduction / max-over-ranks must agree on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, STEPS, SEED = 150, 60, 77
KW = dict(size=10, num_balls=15, agents_index=[3, 5], balls_index=[0, 1, 2], balls_reward=[1, 1, 1], respawn=True,
          layout="quadrants_respawn", time_limit=50)


def _run_shard(base, n):
    import oracle as oc
    o = oc.CollectOracle(oc.make_collect_cfg(**KW), n)
    r = oc.PhiloxRng(seed=SEED, env_id_base=base)
    o.reset(r)
    acts = np.random.default_rng(3).integers(0, 4, size=(STEPS, TOTAL, 2)).astype(np.int8)
    total_reward, obs = 0.0, None
    for t in range(STEPS):
        obs, rew, term, trunc = o.step(acts[t, base:base + n], r, autoreset=True)
        total_reward += float(rew.sum())
    return obs, total_reward, int(o.info.sum())


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_multigrid_b200 import sharding
    assert sharding.world_info() == (rank, rank, world)
    base, n = sharding.shard_range(TOTAL, rank, world)
    obs, rsum, pick = _run_shard(base, n)
    stats = sharding.reduce_stats({"env_steps": n * STEPS, "reward": rsum, "pickups": pick})
    slowest = sharding.max_over_ranks(float(rank + 1))
    gathered = [None] * world
    dist.all_gather_object(gathered, (base, n, obs))
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), obs=np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])]),
                 env_steps=stats["env_steps"], reward=stats["reward"], pickups=stats["pickups"], slowest=slowest)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions():
    from gym_multigrid_b200.sharding import shard_range
    for total, world in ((150, 2), (65536, 8), (10, 3), (7, 8)):
        parts = [shard_range(total, r, world) for r in range(world)]
        assert parts[0][0] == 0 and sum(n for _, n in parts) == total
        for (b0, n0), (b1, _) in zip(parts, parts[1:]):
            assert b0 + n0 == b1
        assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_rank_gloo_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "out.npz")
    obs_full, rsum_full, pick_full = _run_shard(0, TOTAL)
    assert np.array_equal(z["obs"], obs_full)
    assert float(z["reward"]) == rsum_full and int(z["pickups"]) == pick_full
    assert int(z["env_steps"]) == TOTAL * STEPS and float(z["slowest"]) == 2.0
