"""Batched, burst-sharded scoring loop of the SyntheticBurst validation protocol.

The reference's `compute_score` (evaluation/synburst/compute_score.py:35-126) walks the validation set one burst at a time:
forward at batch 1, 14-bit quantisation `(pred.clamp(0, 1) * 2 ** 14).short().float() / 2 ** 14` (:110-111), then each metric
with a `.cpu().item()` per image (:114) -- a host synchronisation per metric per burst.  `score_dataset` keeps the protocol
(same quantisation, same per-image metrics `PSNR(boundary_ignore=40)` / `SSIM(boundary_ignore=40, use_for_loss=False)`
averaged over the images, :47-60,122) and changes the schedule: bursts are stacked into batches, every rank scores a
contiguous shard of the set (`sharding.shard_range`), the network writes the int16 form directly from the predictor
epilogue (`net.output_int16`), the metrics come from the fused kernels without leaving the device, and the ranks exchange
one all-reduce of `[sums | counts]` at the end (`sharding.reduce_metric_means`) -- the only host read is the final report.
LPIPS (a pretrained AlexNet from the `lpips` package) is not on this path and is refused.  `saved_dir` reproduces the
`load_saved` branch (:78-88, 100-104): when the directory holds one `<burst_name>.png` per burst (written by
`save_results.save_results` here or by the reference's own script), the predictions are read from the files instead of
running the network.  `dataset` is any indexable of `(burst [N, 4, H, W], gt [3, 8H, 8W], meta_info)` items, the contract of
`SyntheticBurstVal.__getitem__` (dataset/synthetic_burst_val_set.py:38-55).  `compute_score(setting_name, load_saved)` is the
reference's driver on top of it: experiment file -> list of `NetworkParam` -> one row of the report per network."""
from __future__ import annotations

import importlib
import os
from typing import Dict, Optional, Sequence

import torch
import torch.distributed as dist

from ... import sharding
from ...models.loss import msssim
from ...models.loss.image_quality_v2 import PSNR


class TensorBurstSet:
    """In-memory stand-in with the item contract of `SyntheticBurstVal` (dataset/synthetic_burst_val_set.py:38-55)."""

    def __init__(self, bursts: torch.Tensor, gts: torch.Tensor):
        assert bursts.dim() == 5 and gts.dim() == 4 and bursts.shape[0] == gts.shape[0]
        self.bursts, self.gts = bursts, gts

    def __len__(self):
        return self.bursts.shape[0]

    def __getitem__(self, index):
        return self.bursts[index], self.gts[index], {'burst_name': '{:04d}'.format(index)}

    def batch(self, start: int, stop: int):
        """contiguous items as two views (no per-item stacking); pinned storage makes the upload asynchronous"""
        return self.bursts[start:stop], self.gts[start:stop]


class _Stager:
    """Stacks dataset items into two alternating pinned host buffers, so that the upload of batch i is asynchronous and the
    host assembles batch i + 1 while the device works on batch i."""

    def __init__(self):
        self.slots = [{}, {}]
        self.i = 0

    def __call__(self, items, device):
        slot = self.slots[self.i % 2]
        self.i += 1
        if 'event' in slot:
            slot['event'].synchronize()              # the upload that last read this slot has finished
        out = []
        for k in (0, 1):
            first = items[0][k]
            shape = (len(items),) + tuple(first.shape)
            buf = slot.get(k)
            if buf is None or buf.shape[1:] != shape[1:] or buf.shape[0] < shape[0] or buf.dtype != first.dtype:
                buf = slot[k] = torch.empty(shape, dtype=first.dtype).pin_memory()
            torch.stack([it[k] for it in items], out=buf[:shape[0]])
            out.append(buf[:shape[0]].to(device, non_blocking=True))
        slot['event'] = torch.cuda.Event()
        slot['event'].record()
        return out


def dequantize_q14(pred_q: torch.Tensor) -> torch.Tensor:
    """int16 -> float, `net_pred_int.float() / (2 ** 14)` (compute_score.py:111); exact"""
    return pred_q.float() / 2 ** 14


@torch.no_grad()
def score_dataset(net, dataset, metrics: Sequence[str] = ('psnr', 'ssim'), boundary_ignore: int = 40, batch_size: int = 32,
                  device='cuda', burst_sz=None, shard: bool = True, saved_dir: Optional[str] = None) -> Dict[str, float]:
    """Mean per-image metrics of `net` over `dataset` (all ranks' shards), plus 'count' and 'using_saved_results'.
    `net`: a `DBSRNet` of this package (may be None when `saved_dir` holds a complete set of saved predictions)."""
    for m in metrics:
        if m not in ('psnr', 'ssim'):
            raise NotImplementedError(f'metric {m!r} is not provided (psnr / ssim; lpips needs the `lpips` package)')
    from .save_results import load_prediction, saved_results_complete
    using_saved = saved_dir is not None and saved_results_complete(saved_dir, dataset)
    if not using_saved and net is None:
        raise ValueError('no network given and no complete set of saved results to read')
    # shard=False: this rank scores the whole set on its own (no collective), e.g. to cross-check a sharded run
    distributed = shard and dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    lo, hi = sharding.shard_range(len(dataset), rank, world)
    device = torch.device(device)
    psnr_fn = PSNR(boundary_ignore=boundary_ignore)
    was_q = getattr(net, 'output_int16', False)
    if not using_saved:
        net.output_int16 = True
    per_image = []
    stage = _Stager()
    try:
        for start in range(lo, hi, batch_size):
            stop = min(start + batch_size, hi)
            if using_saved:
                items = [dataset[i] for i in range(start, stop)]
                gt = torch.stack([it[1] for it in items]).to(device).float().contiguous()
                pred = torch.cat([load_prediction(os.path.join(saved_dir, it[2]['burst_name'] + '.png'), device) for it in items])
                burst = None
            elif hasattr(dataset, 'batch'):
                burst, gt = (t.to(device, non_blocking=True) for t in dataset.batch(start, stop))
            else:
                burst, gt = stage([dataset[i] for i in range(start, stop)], device)
            if not using_saved:
                gt = gt.float().contiguous()
                if burst_sz is not None:
                    burst = burst[:, :burst_sz]
                pred_q, _ = net(burst.float().contiguous())
                pred = dequantize_q14(pred_q)
            cols = []
            for m in metrics:
                if m == 'psnr':
                    cols.append(psnr_fn.psnr_per_image(pred, gt))
                else:   # image_quality_v2.SSIM(boundary_ignore, use_for_loss=False) per image: 11-tap window, data-derived range
                    cols.append(msssim._stats(pred, gt, 11, None, None, crop=boundary_ignore or 0, fixed_window=True)[0][:, 0])
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


def generate_formatted_report(scores_all: Dict[str, Dict[str, float]], table_name: str = '') -> str:
    """Text table `name | metric ... |` with three decimals (evaluation/common_utils/display_utils.py:15-41)."""
    name_width = max([len(d) for d in scores_all] + [len(table_name)]) + 5
    names = [k for k in next(iter(scores_all.values())) if k != 'count']
    widths = [max(10, len(k) + 3) for k in names]
    text = '\n{: <{w}} |'.format(table_name, w=name_width)
    for k, w in zip(names, widths):
        text += ' {: <{w}} |'.format(k, w=w)
    text += '\n'
    for net_name, scores in scores_all.items():
        text += '{: <{w}} |'.format(net_name, w=name_width)
        for k, w in zip(names, widths):
            text += ' {: <{w}} |'.format('{:0.3f}'.format(scores[k]), w=w)
        text += '\n'
    return text


def load_experiment(setting_name: str, dataset_name: str = 'synburst'):
    """list of `NetworkParam` of an experiment: `evaluation/<dataset>/experiments/<setting_name>.py::main()` of this package
    (compute_score.py:41-43), or -- a dotted name -- any importable module with a `main()`"""
    mod = setting_name if '.' in setting_name else '{}.evaluation.{}.experiments.{}'.format(__name__.split('.')[0], dataset_name, setting_name)
    return getattr(importlib.import_module(mod), 'main')()


def compute_score(setting_name, load_saved=False, dataset=None, metrics: Sequence[str] = ('psnr', 'ssim'), batch_size: int = 32,
                  device='cuda', verbose: bool = True) -> Dict[str, Dict[str, float]]:
    """The reference's `compute_score(setting_name, load_saved=False)` (evaluation/synburst/compute_score.py:36-122): scores every
    network of the experiment on the SyntheticBurst validation set and prints the report.  Predictions are looked up /
    expected under `<save_data_path>/synburst/<unique_name>`; with `load_saved` a complete set of saved files replaces the
    network run (a `NetworkParam` with only `unique_name` is scored from downloaded predictions that way).  Returns
    {display name: {metric: mean}}.  `dataset` defaults to `SyntheticBurstVal()`; LPIPS is not provided (see `score_dataset`)."""
    from ...admin.environment import env_settings
    from .save_results import saved_results_complete
    if dataset is None:
        from ...dataset.synthetic_burst_val_set import SyntheticBurstVal
        dataset = SyntheticBurstVal()
    base_results_dir = env_settings().save_data_path
    scores_all = {}
    for n in load_experiment(setting_name, 'synburst'):
        out_dir = '{}/synburst/{}'.format(base_results_dir, n.get_unique_name())
        using_saved = bool(load_saved) and saved_results_complete(out_dir, dataset)
        net = None
        if not using_saved:
            net = n.load_net()
            net.to(device).train(False)
        s = score_dataset(net, dataset, metrics=metrics, boundary_ignore=40, batch_size=batch_size, device=device,
                          burst_sz=n.burst_sz, saved_dir=out_dir if using_saved else None)
        scores_all[n.get_display_name()] = {m: s[m] for m in metrics}
    if verbose:
        print(generate_formatted_report(scores_all))
    return scores_all


if __name__ == '__main__':
    import argparse
    parser = argparse.ArgumentParser(description='Compute scores on the SyntheticBurst validation set. With --load_saved, '
                                                 'saved predictions are used whenever a complete set exists.')
    parser.add_argument('setting', type=str, help='Name of experiment setting')
    parser.add_argument('--load_saved', dest='load_saved', action='store_true', default=False)
    args = parser.parse_args()
    compute_score(args.setting, args.load_saved)
