"""Multi-GPU plumbing for the workloads that shard: one process per GPU over
torch.distributed.  Single-scan registration does not shard (replicas only); independent sequences /
LiDAR streams are dealt round-robin to ranks with no data-path collective; the loop-closure descriptor
search keeps a shard of the database per rank and merges per-rank top-k candidates with one all_gather.
Nothing here touches point data: it is host logic, testable with the gloo backend on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def assign(n_units: int, world: int, rank: int) -> list[int]:
    """Units (sequences, LiDAR streams, database shards) of `rank`: unit u lives on rank u mod world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_units, world))


def shard_bounds(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of a database of n_items for `rank` (ids keep their global meaning)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device="cpu") -> list[float]:
    """Element-wise MAX over ranks (timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def sum_over_ranks(values, device="cpu") -> list[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.cpu()]


def merge_topk(local_dist: torch.Tensor, local_id: torch.Tensor, k: int):
    """Global top-k (smallest distance, ties by smaller id) from per-rank candidates.

    local_dist / local_id: (q, k_local) on every rank, ids global.  Every rank gets the same (q, k) result:
    one all_gather of k_local * 12 bytes per query per rank, then an identical merge everywhere — the
    ring-key stage of the reference's loop search (SceneRecognitionScanContext.hpp:267-279, nanoflann
    k = 10), done on database shards.
    """
    if local_dist.shape != local_id.shape or local_dist.dim() != 2:
        raise ValueError("expected (q, k_local) tensors of equal shape")
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        ds = [torch.empty_like(local_dist) for _ in range(world)]
        ids = [torch.empty_like(local_id) for _ in range(world)]
        dist.all_gather(ds, local_dist.contiguous())
        dist.all_gather(ids, local_id.contiguous())
        d = torch.cat(ds, dim=1)
        i = torch.cat(ids, dim=1)
    else:
        d, i = local_dist, local_id
    # lexicographic (distance, id): sort by id first, then a stable sort by distance
    o1 = torch.argsort(i, dim=1, stable=True)
    d1, i1 = torch.gather(d, 1, o1), torch.gather(i, 1, o1)
    o2 = torch.argsort(d1, dim=1, stable=True)
    d2, i2 = torch.gather(d1, 1, o2), torch.gather(i1, 1, o2)
    return d2[:, :k].contiguous(), i2[:, :k].contiguous()
