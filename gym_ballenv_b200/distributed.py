"""Multi-GPU plumbing: one process per GPU, environments partitioned by global id, no data-path collective.

The reference is single-process (one env object per script, e.g. examples/ball_cnn_ac3.py:468); environments never
interact (gym_ballenv/envs/ballenv_env.py:232-289), so a job of ``total_envs`` environments shards into contiguous
ranges of global ids.  Every random draw is addressed by the *global* id (csrc/ballenv_rng.cuh), hence a given
environment has the same trajectory on 1, 2, 4 or 8 GPUs.  The only thing that crosses GPUs is the 16-double
episode-statistics vector (``BallVecEnv.stats_tensor``), summed with one small all-reduce (NCCL over NVLink on
CUDA tensors; gloo on CPU tensors in the tests).

This module has no CUDA dependency of its own, so the sharding arithmetic is testable on CPU.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

NUM_STATS = 16


def shard_bounds(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(global offset, count) of the contiguous id range owned by ``rank``: ranges differ by at most one
    environment, cover [0, total_envs) exactly and are ordered by rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %d not in [0, %d)" % (rank, world))
    if total_envs < 0:
        raise ValueError("total_envs < 0")
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def rank_world() -> Tuple[int, int, int]:
    """(rank, world size, local rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def make_sharded_env(total_envs: int, rank: Optional[int] = None, world: Optional[int] = None, device=None, **kw):
    """This rank's ``BallVecEnv`` of a ``total_envs`` job (keyword arguments as for BallVecEnv)."""
    import torch
    from .vec_env import BallVecEnv
    r, w, local = rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    offset, count = shard_bounds(total_envs, rank, world)
    if count == 0:
        raise ValueError("rank %d owns no environments (total_envs=%d, world=%d)" % (rank, total_envs, world))
    if device is None:
        device = torch.device("cuda", local)
    return BallVecEnv(count, device=device, global_env_offset=offset, **kw)


def allreduce_stats(stats, group=None, async_op: bool = False):
    """Sum the episode-statistics vector over all ranks, in place on a copy -> (tensor, work or None).

    ``stats`` is ``BallVecEnv.stats_tensor`` (float64 [16], device) or any float64 [16] tensor; the env's own
    vector is not modified (each rank keeps counting locally)."""
    import torch
    import torch.distributed as dist
    buf = stats.detach().clone()
    if buf.dtype != torch.float64 or buf.numel() != NUM_STATS:
        raise ValueError("stats must be float64 [%d]" % NUM_STATS)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return buf, None
    work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return buf, (work if async_op else None)
