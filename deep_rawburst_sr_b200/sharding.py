"""Burst sharding across ranks.  Bursts are independent, so the data path needs no collective: every
rank takes a contiguous chunk of the batch dimension, runs the forward on its own GPU, and -- only if the caller wants
the full batch in one place -- outputs / metric scalars are gathered with one NCCL collective after the forward.
The reference analogue is `MultiGPU(nn.DataParallel)` scattering dim 0 (admin/multigpu.py:8-14)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [start, stop) of `total` bursts owned by `rank`; the first `total % world` ranks get one extra"""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_bursts(local: torch.Tensor, total: int) -> torch.Tensor:
    """all-gather per-rank outputs [b_local, ...] (ragged over ranks) into [total, ...] in burst order"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = shard_sizes(total, world)
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def sharded_forward(forward: Callable[[torch.Tensor], torch.Tensor], bursts: torch.Tensor, gather: bool = True):
    """run `forward` on this rank's shard of `bursts` [B, N, 4, H, W]; optionally gather the predictions"""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(bursts.shape[0], rank, world)
    pred = forward(bursts[lo:hi])
    return gather_bursts(pred, bursts.shape[0]) if gather else pred


def max_over_ranks(value: float, device) -> float:
    """device-time reduction used by bench.py: the step time of a multi-GPU run is the slowest rank's"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
