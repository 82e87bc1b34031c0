"""Env-index sharding: the batch splits over GPUs (and over streams inside one GPU) in contiguous
blocks with no exchange step — environments are independent, so there is no collective on the
path."""
from __future__ import annotations

from typing import List, Tuple


def env_shard(n_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the env indices owned by `rank` of `world`: contiguous, balanced to within one
    env, in rank order (so that concatenating the shards restores the batch order)."""
    if not (0 <= rank < world) or n_envs < 0:
        raise ValueError((n_envs, rank, world))
    base, extra = divmod(n_envs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_shards(n_envs: int, world: int) -> List[Tuple[int, int]]:
    return [env_shard(n_envs, r, world) for r in range(world)]
