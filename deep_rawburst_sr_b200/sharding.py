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
    dist.all_gather([o.view(torch.uint8) for o in out], buf.view(torch.uint8))      # byte views: NCCL has no int16
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


def reduce_metric_means(per_image: torch.Tensor, collective: bool = True) -> torch.Tensor:
    """Mean over ALL ranks' images of per-image metric values [b_local] or [b_local, k] (e.g. the outputs of
    `PSNR.psnr_per_image` / `msssim.ssim(size_average=False)`), with one all_reduce(SUM) of `[sums | counts]` instead of a
    gather of the images ("plus one all_reduce(SUM) of [sum psnr, count]").  Non-finite values are dropped per
    metric, as the reference's PSNR does (models/loss/image_quality_v2.py:97); a metric with no finite value yields 0.
    Stays on the tensor's device (fp64 accumulation), no host synchronisation."""
    v = per_image if per_image.dim() == 2 else per_image.unsqueeze(1)
    ok = torch.isfinite(v)
    acc = torch.cat([torch.where(ok, v, torch.zeros_like(v)).double().sum(0), ok.double().sum(0)])
    if collective and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    k = v.shape[1]
    mean = (acc[:k] / acc[k:].clamp(min=1)).to(per_image.dtype)
    return mean if per_image.dim() == 2 else mean[0]


class OutputGatherer:
    """Overlapped gather of per-rank predictions ("run it on a side stream overlapped with the next batch").

    `submit(local)` copies this rank's `pred` shard (possibly a static CUDA-graph output that the next forward overwrites)
    into one of `depth` staging slots on the CURRENT stream, then all-gathers the slot into `[total, ...]` on a side
    stream over NCCL / NVLink while the caller launches the next forward.  It returns `(gathered, event)`; `gathered` is
    valid once `event` has completed and is reused `depth` submissions later.  Ragged shards are padded to the largest
    shard (the first `total % world` ranks own one burst more, `shard_range`).  On CPU tensors (gloo, tests) the same
    logic runs synchronously without streams."""

    def __init__(self, total: int, depth: int = 2):
        assert depth >= 1
        self.total, self.depth = total, depth
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.sizes = shard_sizes(total, self.world)
        self.pad = max(self.sizes)
        self._stage = [None] * depth
        self._out = [None] * depth
        self._done = [None] * depth
        self._side = None
        self._i = 0

    def submit(self, local: torch.Tensor):
        k = self._i % self.depth
        self._i += 1
        cuda = local.is_cuda
        tail = tuple(local.shape[1:])
        if self._stage[k] is None or tuple(self._stage[k].shape[1:]) != tail or self._stage[k].dtype != local.dtype:
            self._stage[k] = local.new_zeros((self.pad,) + tail)
            self._out[k] = local.new_empty((self.world * self.pad,) + tail)
            self._done[k] = None
        cur = torch.cuda.current_stream(local.device) if cuda else None
        if cuda and self._done[k] is not None:
            cur.wait_event(self._done[k])            # the gather that last read this slot has finished
        self._stage[k][:local.shape[0]].copy_(local, non_blocking=True)
        if self.world == 1:
            ev = None
            if cuda:
                ev = torch.cuda.Event()
                ev.record(cur)
            return self._stage[k][:self.total], ev
        if not cuda:
            dist.all_gather_into_tensor(self._bytes(self._out[k]), self._bytes(self._stage[k]))
            return self._compact(self._out[k]), None
        if self._side is None:
            self._side = torch.cuda.Stream(device=local.device)
        staged = torch.cuda.Event()
        staged.record(cur)
        self._side.wait_event(staged)
        with torch.cuda.stream(self._side):
            dist.all_gather_into_tensor(self._bytes(self._out[k]), self._bytes(self._stage[k]))
            # ragged shards: the compaction READS the gathered buffer, so it must be ordered after the collective -- on the
            # side stream, before the event the caller waits on (on the current stream it would race with the gather)
            gathered = self._compact(self._out[k])
            ev = torch.cuda.Event()
            ev.record(self._side)
        if gathered is not self._out[k]:
            gathered.record_stream(cur)               # allocated on the side stream, consumed on the caller's
        self._done[k] = ev
        return gathered, ev

    @staticmethod
    def _bytes(t: torch.Tensor) -> torch.Tensor:
        """the gather is pure data movement: hand the collective a byte view, so that every element type works (NCCL has no
        int16, the 14-bit quantised prediction format)"""
        return t.view(torch.uint8)

    def _compact(self, out: torch.Tensor) -> torch.Tensor:
        if all(s == self.pad for s in self.sizes):
            return out                                # even shards: the gathered buffer is already [total, ...]
        return torch.cat([out[r * self.pad:r * self.pad + s] for r, s in enumerate(self.sizes)], dim=0)
