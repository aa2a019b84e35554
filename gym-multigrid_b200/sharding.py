"""Multi-GPU plumbing: envs shard by contiguous global-id ranges, one process per GPU.

There is NO collective on the step path (envs are independent; Philox streams are keyed by the
*global* env id, so results do not depend on how many GPUs the envs are spread over).  The only
communication is the optional end-of-run reduction of a few scalars (`reduce_stats`) and the
barrier / max-over-ranks around timed regions, through torch.distributed (NCCL on GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

import os


def world_info():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous, balanced partition of [0, total_envs): returns (env_id_base, num_envs) of `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    q, r = divmod(int(total_envs), int(world_size))
    base = rank * q + min(rank, r)
    return base, q + (1 if rank < r else 0)


def make_vec_sharded(env_id: str, total_envs: int, seed: int = 0, **kwargs):
    """This rank's shard of a `total_envs`-env job on its local GPU (cuda:LOCAL_RANK)."""
    from . import make_vec
    rank, local_rank, world = world_info()
    base, n = shard_range(total_envs, rank, world)
    return make_vec(env_id, n, device=f"cuda:{local_rank}", seed=seed, env_id_base=base, **kwargs)


def reduce_stats(stats: dict, device=None) -> dict:
    """Sum a dict of python/torch scalars over all ranks (end-of-run statistics: total env-steps,
    reward sums, pickup counters).  A no-op outside torch.distributed."""
    import torch
    import torch.distributed as dist
    keys = sorted(stats)
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
