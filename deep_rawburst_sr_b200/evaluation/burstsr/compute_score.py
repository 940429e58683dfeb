"""Batched, burst-sharded scoring loop of the BurstSR (real-world) validation protocol (SURVEY.md 8(f) ranks 1-3 together).

The reference's `compute_score` (evaluation/burstsr/compute_score.py:38-134) scores one burst at a time: forward, 14-bit
quantisation (:121-122), `SpatialColorAlignment(pwcnet, sr_factor=4)(pred, gt, burst)` (:72-73,124-125: PWC-Net at the output
resolution, two warps, per-image colour matrix, validity mask), then every metric with the mask and a `.cpu().item()` (:126-128).
`score_dataset` keeps that protocol per image -- `PSNR(boundary_ignore=40)` / `SSIM(boundary_ignore=40, use_for_loss=False)`
with `valid`, each image normalised by its own maximum before the alignment net (`per_image_norm`, what batch 1 does) -- and
changes the schedule exactly as `evaluation/synburst/compute_score.py` of this package does: batches, a contiguous shard of
the set per rank, int16 predictions from the predictor epilogue, masked metrics from the fused kernels, one all-reduce of
`[sums | counts]`, one host read.  Items are dicts with 'burst' [N, 4, H, W], 'frame_gt' [3, 8H, 8W] and 'burst_name', the
contract of `BurstSRDataset.__getitem__` as the reference loop reads it (:99-102).  LPIPS is refused (not on this path)."""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.distributed as dist

from ... import sharding
from ...models.loss import msssim
from ...models.loss.image_quality_v2 import PSNR
from ...models.loss.spatial_color_alignment import SpatialColorAlignment
from ..synburst.compute_score import dequantize_q14, generate_formatted_report  # noqa: F401


@torch.no_grad()
def score_dataset(net, dataset, alignment_net, metrics: Sequence[str] = ('psnr', 'ssim'), boundary_ignore: int = 40,
                  batch_size: int = 16, device='cuda', burst_sz=None, sr_factor: int = 4, shard: bool = True) -> Dict[str, float]:
    """Mean per-image aligned metrics of `net` over `dataset` (all ranks' shards), plus 'count'."""
    for m in metrics:
        if m not in ('psnr', 'ssim'):
            raise NotImplementedError(f'metric {m!r} is not provided (psnr / ssim; lpips needs the `lpips` package)')
    # shard=False: this rank scores the whole set on its own (no collective), e.g. to cross-check a sharded run
    distributed = shard and dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    lo, hi = sharding.shard_range(len(dataset), rank, world)
    device = torch.device(device)
    sca = SpatialColorAlignment(alignment_net.eval(), sr_factor=sr_factor)
    sca.to(device)
    sca.per_image_norm = True
    psnr_fn = PSNR(boundary_ignore=boundary_ignore)
    was_q = getattr(net, 'output_int16', False)
    net.output_int16 = True
    per_image = []
    try:
        for start in range(lo, hi, batch_size):
            items = [dataset[i] for i in range(start, min(start + batch_size, hi))]
            burst = torch.stack([it['burst'] for it in items]).to(device, non_blocking=True).float().contiguous()
            gt = torch.stack([it['frame_gt'] for it in items]).to(device, non_blocking=True).float().contiguous()
            if burst_sz is not None:
                burst = burst[:, :burst_sz].contiguous()
            pred_q, _ = net(burst)
            pred = dequantize_q14(pred_q)
            pred_m, valid = sca(pred, gt, burst)
            cols = []
            for m in metrics:
                if m == 'psnr':
                    cols.append(psnr_fn.psnr_per_image(pred_m, gt, valid))
                else:   # image_quality_v2.SSIM(boundary_ignore, use_for_loss=False)(pred_m, gt, valid) per image
                    st, _ = msssim._stats(pred_m, gt, 11, None, None, crop=boundary_ignore or 0, fixed_window=True, valid=valid)
                    cols.append(st[:, 0] / (st[:, 1] + 1e-12))
            per_image.append(torch.stack(cols, dim=1))
    finally:
        net.output_int16 = was_q
    local = torch.cat(per_image) if per_image else torch.zeros(0, len(metrics), device=device)
    mean = sharding.reduce_metric_means(local, collective=distributed)
    out = {m: float(v) for m, v in zip(metrics, mean.cpu())}
    out['count'] = len(dataset)
    return out
