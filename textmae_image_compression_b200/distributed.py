"""Data-parallel plumbing for the path: images are sharded across ranks, weights are replicated, and the only
exchange is one all-reduce(SUM) of {sum log2 likelihood, pixels} (16 bytes) to form the dataset-level rate -
the same aggregate the reference forms with `all_reduce_mean` / `MetricLogger.synchronize_between_processes`
(/root/reference/models/Compression/common/distributed.py:25-33, common/logger.py:29-40).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_round_robin(n_items: int, rank: int, world: int):
    """Tile indices owned by `rank` when the tiles of one large image are dealt round-robin."""
    return list(range(rank, n_items, world))


def aggregate_rate(rate_sums: torch.Tensor) -> torch.Tensor:
    """rate_sums: f64 [2] = {sum log2 likelihood, pixels} of the local shard.  Returns the global bits-per-pixel
    (0-d f64 tensor on the same device).  One 16-byte all-reduce; identity when no process group exists."""
    t = rate_sums.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return -t[0] / t[1]


def gather_per_image(bpp_local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Optional: all-gather the per-image bpp of every shard (ragged shards allowed) in global image order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return bpp_local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    maxlen = max(e - b for b, e in sizes)
    buf = torch.zeros(maxlen, dtype=bpp_local.dtype, device=bpp_local.device)
    buf[: bpp_local.numel()] = bpp_local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return torch.cat([o[: e - b] for o, (b, e) in zip(outs, sizes)])
