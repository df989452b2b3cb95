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


class AsyncStats:
    """Periodic whole-job episode statistics without touching the step's critical path: every `every`
    ticks the local int64[8] counters are read and summed over all ranks on a SIDE stream through a process group of
    their own (its own NCCL communicator, so the reduction never queues behind -- or in front of -- a collective the
    training loop issues on the default group); `latest()` hands out the last reduction that has completed, never
    blocking.  The counters are monotone, so a reduction that races with the next steps' increments is a valid
    snapshot "as of about" its tick; `ShardedRunner.stats()` is the exact, blocking read.
    read_local: callable(stream or None) -> the local int64[8] tensor (on the group's device)."""

    def __init__(self, read_local, every=16, device=None):
        import torch
        import torch.distributed as dist
        self.read_local, self.every, self.ticks = read_local, int(every), 0
        self.dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.group = dist.new_group() if self.dist_on else None
        self.cuda = device is not None and torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.pending = []                      # (tick, tensor, event or work)
        self.done = None                       # (tick, tensor)
        self.reductions = 0

    def tick(self):
        """Call once per env step (all ranks, same cadence)."""
        import torch
        import torch.distributed as dist
        self.ticks += 1
        if self.ticks % self.every:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())          # the counters of the steps issued so far
            with torch.cuda.stream(self.stream):
                buf = self.read_local(self.stream).clone()
                if self.dist_on:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.pending.append((self.ticks, buf, ev))
        else:
            buf = self.read_local(None).clone()
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if self.dist_on else None
            self.pending.append((self.ticks, buf, work))
        self.reductions += 1

    def _finished(self, handle):
        if handle is None:
            return True
        return handle.query() if hasattr(handle, "query") else handle.is_completed()

    def latest(self, wait=False):
        """(tick, {name: value}) of the newest completed reduction, or None.  wait=True blocks for the newest issued one."""
        while self.pending and (wait or self._finished(self.pending[0][2])):
            tick, buf, handle = self.pending.pop(0)
            if wait and handle is not None:
                handle.synchronize() if hasattr(handle, "synchronize") else handle.wait()
            self.done = (tick, buf)
        if self.done is None:
            return None
        return self.done[0], dict(zip(abi.STAT_NAMES, (int(v) for v in self.done[1].cpu().tolist())))


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
        self._async = None
        self._puzzle_idx = torch.from_numpy(
            shard_puzzles(n_global, self.rank, self.world, self.env.store.n_puzzles)).to(self.device)

    def reset(self):
        return self.env.reset(puzzle_idx=self._puzzle_idx)

    def step(self, actions):
        out = self.env.step(actions)
        if self._async is not None:
            self._async.tick()
        return out

    def stats_async(self, every=16):
        """Reduce the episode statistics over all ranks every `every` steps on a side stream (AsyncStats); read the
        newest completed result with `latest_stats()`."""
        import ctypes as C
        import torch
        env = self.env
        buf = torch.zeros(abi.N_STATS, dtype=torch.int64, device=self.device)

        def read_local(stream):
            with torch.cuda.device(self.device):
                abi.check(env.lib.sparc_stats(env.handle, buf.data_ptr(), 0, C.c_void_p(stream.cuda_stream)))
            return buf
        self._async = AsyncStats(read_local, every=every, device=self.device)
        return self._async

    def latest_stats(self, wait=False):
        return self._async.latest(wait) if self._async is not None else None

    def stats(self, clear=False):
        """Whole-job episode statistics (all ranks must call)."""
        return reduce_stats(self.env.stats(clear=clear))
