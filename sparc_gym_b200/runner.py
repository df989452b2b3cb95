"""Env-sharded multi-GPU runner: one process per GPU, each with its own shard of the envs and a
replica of the puzzle store; no data-path collective.  The only cross-GPU traffic is the
all-reduce of the int64[8] episode statistics (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The reference has no counterpart (single process, single env); the sharding keeps its
puzzle-cursor convention: GLOBAL env i starts on puzzle i mod pool (cf. SG.py:80, 1087).
"""
from __future__ import annotations

import os

import numpy as np

from . import abi


def shard_range(n_global, rank, world):
    """Contiguous env range [lo, hi) of `rank`; the first n_global % world ranks get one more."""
    base, rem = divmod(int(n_global), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_puzzles(n_global, rank, world, n_puzzles):
    """Initial puzzle index of every local env: global env i -> puzzle i mod pool."""
    lo, hi = shard_range(n_global, rank, world)
    return (np.arange(lo, hi, dtype=np.int64) % n_puzzles).astype(np.int32)


def reduce_stats(local_stats, group=None):
    """Sum the int64[8] statistics over all ranks (exact: integer fixed point).  `local_stats` is a
    torch tensor on the device of the process group's backend; returns a dict by name."""
    import torch.distributed as dist
    t = local_stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(abi.STAT_NAMES, (int(v) for v in t.cpu().tolist())))


class ShardedRunner:
    """`n_global` envs split over WORLD_SIZE ranks (torchrun environment); rank r owns
    shard_range(n_global, r, world) on GPU LOCAL_RANK."""

    def __init__(self, puzzles, n_global, rank=None, world=None, local_rank=None, **env_kwargs):
        import torch
        from .vec_env import SparcVecEnv
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0)) if local_rank is None else local_rank
        self.n_global = int(n_global)
        self.lo, self.hi = shard_range(n_global, self.rank, self.world)
        self.device = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.device)
        self.env = SparcVecEnv(puzzles, self.hi - self.lo, device=self.device, **env_kwargs)
        self._puzzle_idx = torch.from_numpy(
            shard_puzzles(n_global, self.rank, self.world, self.env.store.n_puzzles)).to(self.device)

    def reset(self):
        return self.env.reset(puzzle_idx=self._puzzle_idx)

    def step(self, actions):
        return self.env.step(actions)

    def stats(self, clear=False):
        """Whole-job episode statistics (all ranks must call)."""
        return reduce_stats(self.env.stats(clear=clear))
