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

import os
from typing import Dict, Optional, Sequence

import torch
import torch.distributed as dist

from ... import sharding
from ...models.loss import msssim
from ...models.loss.image_quality_v2 import PSNR
from ...models.loss.spatial_color_alignment import SpatialColorAlignment
from ..synburst.compute_score import dequantize_q14, generate_formatted_report  # noqa: F401
from ..synburst.save_results import load_prediction


@torch.no_grad()
def score_dataset(net, dataset, alignment_net, metrics: Sequence[str] = ('psnr', 'ssim'), boundary_ignore: int = 40,
                  batch_size: int = 16, device='cuda', burst_sz=None, sr_factor: int = 4, shard: bool = True,
                  saved_dir: Optional[str] = None) -> Dict[str, float]:
    """Mean per-image aligned metrics of `net` over `dataset` (all ranks' shards), plus 'count' and 'using_saved_results'.
    `saved_dir`: the `load_saved` branch (compute_score.py:84-93, 112-115) -- with at least one PNG per burst the predictions
    are read from `<saved_dir>/<burst_name>.png` instead of running `net` (which may then be None)."""
    for m in metrics:
        if m not in ('psnr', 'ssim'):
            raise NotImplementedError(f'metric {m!r} is not provided (psnr / ssim; lpips needs the `lpips` package)')
    # shard=False: this rank scores the whole set on its own (no collective), e.g. to cross-check a sharded run
    distributed = shard and dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    lo, hi = sharding.shard_range(len(dataset), rank, world)
    device = torch.device(device)
    using_saved = saved_dir is not None and os.path.isdir(saved_dir) and \
        len([r for r in os.listdir(saved_dir) if r[-3:] == 'png']) >= len(dataset)
    if not using_saved and net is None:
        raise ValueError('no network given and no complete set of saved results to read')
    sca = SpatialColorAlignment(alignment_net.eval(), sr_factor=sr_factor)
    sca.to(device)
    sca.per_image_norm = True
    psnr_fn = PSNR(boundary_ignore=boundary_ignore)
    was_q = getattr(net, 'output_int16', False)
    if not using_saved:
        net.output_int16 = True
    per_image = []
    try:
        for start in range(lo, hi, batch_size):
            items = [dataset[i] for i in range(start, min(start + batch_size, hi))]
            burst = torch.stack([it['burst'] for it in items]).to(device, non_blocking=True).float().contiguous()
            gt = torch.stack([it['frame_gt'] for it in items]).to(device, non_blocking=True).float().contiguous()
            if burst_sz is not None:
                burst = burst[:, :burst_sz].contiguous()
            if using_saved:
                pred = torch.cat([load_prediction(os.path.join(saved_dir, it['burst_name'] + '.png'), device) for it in items])
            else:
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
        if not using_saved:
            net.output_int16 = was_q
    local = torch.cat(per_image) if per_image else torch.zeros(0, len(metrics), device=device)
    mean = sharding.reduce_metric_means(local, collective=distributed)
    out = {m: float(v) for m, v in zip(metrics, mean.cpu())}
    out['count'] = len(dataset)
    out['using_saved_results'] = bool(using_saved)
    return out


def compute_score(setting_name, load_saved=False, dataset=None, alignment_net=None, metrics: Sequence[str] = ('psnr', 'ssim'),
                  batch_size: int = 16, device='cuda', verbose: bool = True) -> Dict[str, Dict[str, float]]:
    """The reference's `compute_score(setting_name, load_saved=False)` for BurstSR (evaluation/burstsr/compute_score.py:38-136):
    every network of `evaluation/burstsr/experiments/<setting_name>.py`, predictions under `<save_data_path>/burstsr/<unique_name>`,
    aligned PSNR / SSIM with `boundary_ignore=40`, one report.  `alignment_net` defaults to the reference's choice,
    `PWCNet(load_pretrained=True, weights_path='<pretrained_nets_dir>/pwcnet-network-default.pth')`.  The BurstSR dataset classes
    are out of scope: `dataset` is any indexable of `{'burst', 'frame_gt', 'burst_name'}` items, the contract of
    `get_burstsr_val_set()` (dataset/burstsr_dataset.py:292-302)."""
    from ...admin.environment import env_settings
    from ...models.alignment.pwcnet import PWCNet
    from ..synburst.compute_score import load_experiment
    if dataset is None:
        raise ValueError('compute_score(burstsr): pass dataset= (items with burst / frame_gt / burst_name); the BurstSR RAW '
                         'container classes are not part of this package')
    if alignment_net is None:
        alignment_net = PWCNet(load_pretrained=True,
                               weights_path='{}/pwcnet-network-default.pth'.format(env_settings().pretrained_nets_dir))
    base_results_dir = env_settings().save_data_path
    scores_all = {}
    for n in load_experiment(setting_name, 'burstsr'):
        out_dir = '{}/burstsr/{}'.format(base_results_dir, n.get_unique_name())
        using_saved = bool(load_saved) and os.path.isdir(out_dir) and \
            len([r for r in os.listdir(out_dir) if r[-3:] == 'png']) >= len(dataset)
        net = None
        if not using_saved:
            net = n.load_net()
            net.to(device).train(False)
        s = score_dataset(net, dataset, alignment_net, metrics=metrics, boundary_ignore=40, batch_size=batch_size, device=device,
                          burst_sz=n.burst_sz, saved_dir=out_dir if using_saved else None)
        scores_all[n.get_display_name()] = {m: s[m] for m in metrics}
    if verbose:
        print(generate_formatted_report(scores_all))
    return scores_all
