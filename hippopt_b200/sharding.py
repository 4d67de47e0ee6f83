"""Instance sharding over the GPUs of one box.

Problem instances share nothing, so rank r owns the contiguous block of instances
``shard_range(B, r, world)`` and the data path has no collective; the only exchange is the gather of
per-instance results (objective, constraint violation, ... a few doubles per instance) at the end of
a step, done with ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_instances: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the instances owned by ``rank``; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n_instances, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_instances(local: torch.Tensor, n_instances: int, group=None) -> torch.Tensor:
    """All-gather per-instance rows (first dim = local instances, in shard order) into the global
    (n_instances, ...) tensor on every rank.  Shards may differ in size by one row (padded)."""
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[0] != n_instances:
            raise ValueError("single process must hold every instance")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(n_instances, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} instances, expected {hi - lo}")
    width = -(-n_instances // world)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_instances, r, world)
        parts.append(out[r * width: r * width + (b - a)])
    return torch.cat(parts, dim=0)


def gather_rank_rows(row: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather ONE fixed-length row of per-rank scalars (counts, seconds, ...) into (world, len) on every rank."""
    if not dist.is_available() or not dist.is_initialized():
        return row[None].clone()
    world = dist.get_world_size(group)
    out = torch.empty((world, row.numel()), dtype=row.dtype, device=row.device)
    dist.all_gather_into_tensor(out, row[None].contiguous(), group=group)
    return out


def whole_job_rate(rows: torch.Tensor, count_col: int, seconds_col: int) -> tuple[float, float, float]:
    """Whole-job throughput of work sharded over ranks: (units of all ranks, slowest rank's seconds, units per second).
    The job is finished when its slowest rank is: the rate is the SUM of the units over the MAX of the times -- never a
    sum of per-rank rates."""
    units = float(rows[:, count_col].sum())
    slow = float(rows[:, seconds_col].max())
    return units, slow, (units / slow if slow > 0 else 0.0)
