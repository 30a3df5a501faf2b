"""Multi-GPU plumbing: envs shard by index, one process per GPU, no data-path collective.
The only collective of the hot path is one all-reduce (sum) of the 8-element episode-statistics
vector per rollout — NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(num_envs_total, rank, world_size):
    """Contiguous env-index shard of rank r: start, count  (rank r owns [start, start + count))."""
    per, rem = divmod(int(num_envs_total), int(world_size))
    start = rank * per + min(rank, rem)
    return start, per + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def allreduce_sum(t, group=None):
    """In-place sum over ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def merge_stats(stats_vector, group=None):
    """All-reduce an episode-statistics vector and return it as a dict (see _lib.STAT_NAMES)."""
    from ._lib import STAT_NAMES
    v = allreduce_sum(stats_vector.clone(), group).tolist()
    d = dict(zip(STAT_NAMES, v))
    d["mean_episode_length"] = d["sum_length"] / d["episodes"] if d["episodes"] else float("nan")
    return d
