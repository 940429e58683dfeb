"""Image-quality metrics of the DBSR path with the reference's interface (models/loss/image_quality_v2.py):
`PixelWiseError` (:24-66), `PSNR` (:69-101) and `AlignedL2` (:166-191, the BurstSR loss / metric built on
`SpatialColorAlignment`) and `SSIM` (:104-136, on the fused `dbsr_ssim` kernel of `msssim.py`).  On CUDA fp32 batches
without a `valid` mask `PSNR` takes its per-image MSE from one `dbsr_mse_per_image` launch (boundary_ignore as index
arithmetic) instead of a Python loop of sliced reductions.  LPIPS depends on a package outside this path (`lpips`, a
pretrained AlexNet) and is not provided."""
import math

import torch
import torch.nn as nn

from ... import ops
from . import msssim
from . import spatial_color_alignment as sca_utils


def _interior(b, *tensors):
    """drop `b` boundary pixels on every side of the last two dimensions (boundary_ignore); None entries pass through"""
    if b is None:
        return tensors
    return tuple(t if t is None else t[..., b:-b, b:-b] for t in tensors)


def _masked_mean(err, valid):
    """mean of `err` over the valid pixels; a one-channel mask counts once per channel (image_quality_v2.py:60-64)"""
    mask = valid.float()
    per_mask_elem = err.numel() / valid.numel()
    return (err * mask).sum() / (mask.sum() * per_mask_elem + 1e-12)


_ELEMENTWISE = {'l1': lambda d: d.abs(), 'l2': lambda d: d * d}
_REDUCED = {'l2_sqrt': lambda d: (d * d).sum(dim=-3).sqrt().mean(),
            'charbonnier': lambda d: (d * d + 1e-3 ** 2).sqrt().mean()}


class PixelWiseError(nn.Module):
    """Pixel-wise error ('l1' | 'l2' | 'l2_sqrt' | 'charbonnier'), optionally ignoring `boundary_ignore` boundary pixels and,
    for 'l1' / 'l2', restricted to a validity mask (reference :24-66; there too the other two metrics take no mask)."""
    def __init__(self, metric='l1', boundary_ignore=None):
        super().__init__()
        if metric not in _ELEMENTWISE and metric not in _REDUCED:
            raise Exception
        self.metric = metric
        self.boundary_ignore = boundary_ignore

    def forward(self, pred, gt, valid=None):
        pred, gt, valid = _interior(self.boundary_ignore, pred, gt, valid)
        diff = pred - gt
        if self.metric in _REDUCED:
            if valid is not None:
                raise TypeError(f"metric {self.metric!r} does not take a validity mask")
            return _REDUCED[self.metric](diff)
        err = _ELEMENTWISE[self.metric](diff)
        return err.mean() if valid is None else _masked_mean(err, valid)


class PSNR(nn.Module):
    def __init__(self, boundary_ignore=None, max_value=1.0):
        super().__init__()
        self.l2 = PixelWiseError(metric='l2', boundary_ignore=boundary_ignore)
        self.max_value = max_value

    def psnr(self, pred, gt, valid=None):
        mse = self.l2(pred, gt, valid=valid)
        if getattr(self, 'max_value', 1.0) is not None:
            psnr = 20 * math.log10(getattr(self, 'max_value', 1.0)) - 10.0 * mse.log10()
        else:
            psnr = 20 * gt.max().log10() - 10.0 * mse.log10()
        if torch.isinf(psnr) or torch.isnan(psnr):
            print('invalid psnr')
        return psnr

    def psnr_per_image(self, pred, gt, valid=None):
        """[n] PSNR of every image of a CUDA fp32 batch from one fused launch, no host synchronisation (what a sharded
        evaluation all-reduces instead of gathering images).  valid: optional [n, 1, h, w] bool mask."""
        b = self.l2.boundary_ignore
        mse = ops.mse_per_image(pred.contiguous(), gt.contiguous(), crop=0 if b is None else b, valid=valid)
        return 20 * math.log10(getattr(self, 'max_value', 1.0)) - 10.0 * mse.log10()

    def forward(self, pred, gt, valid=None):
        mask_ok = valid is None or (valid.dim() == 4 and valid.shape[1] == 1 and valid.dtype in (torch.bool, torch.uint8))
        if mask_ok and pred.is_cuda and pred.dim() == 4 and pred.dtype == torch.float32 and gt.dtype == torch.float32 \
                and pred.shape == gt.shape and getattr(self, 'max_value', 1.0) is not None:
            psnr = self.psnr_per_image(pred, gt, valid)
            ok = torch.isfinite(psnr)                      # the reference drops inf / nan images (:97), 0 if none is left
            return torch.where(ok, psnr, torch.zeros_like(psnr)).sum() / ok.sum().clamp(min=1)
        if valid is None:
            psnr_all = [self.psnr(p.unsqueeze(0), g.unsqueeze(0)) for p, g in zip(pred, gt)]
        else:
            psnr_all = [self.psnr(p.unsqueeze(0), g.unsqueeze(0), v.unsqueeze(0)) for p, g, v in zip(pred, gt, valid)]
        psnr_all = [p for p in psnr_all if not (torch.isinf(p) or torch.isnan(p))]
        if len(psnr_all) == 0:
            return 0
        return sum(psnr_all) / len(psnr_all)


class SSIM(nn.Module):
    def __init__(self, boundary_ignore=None, use_for_loss=True):
        super().__init__()
        self.ssim = msssim.SSIM(spatial_out=True)
        self.boundary_ignore = boundary_ignore
        self.use_for_loss = use_for_loss

    def forward(self, pred, gt, valid=None):
        crop = 0 if self.boundary_ignore is None else self.boundary_ignore
        if pred.dim() == 3:
            pred = pred.unsqueeze(0)
            gt = gt.unsqueeze(0)
        if valid is None:
            stats, _ = msssim._stats(pred, gt, self.ssim.window_size, None, None, want_map=False, crop=crop, fixed_window=True)
            loss = stats[:, 0].mean()
        elif valid.dim() == 4 and valid.shape[1] == 1 and valid.dtype in (torch.bool, torch.uint8):
            # masked mean inside the kernel, over the batch: (sum ssim * valid) / (sum valid * C + eps) (:127-131); the kernel
            # crops the mask by index arithmetic exactly like the images (boundary_ignore, then 5 for the 11-tap window)
            stats, _ = msssim._stats(pred, gt, self.ssim.window_size, None, None, crop=crop, fixed_window=True, valid=valid)
            loss = stats[:, 0].sum() / (stats[:, 1].sum() + 1e-12)
        else:                                              # other mask layouts: the map, then the reference's own torch ops
            if crop:
                valid = valid[..., crop:-crop, crop:-crop]
            loss = self.ssim(pred, gt, crop=crop)
            valid = valid[..., 5:-5, 5:-5]  # assume window size 11
            eps = 1e-12
            elem_ratio = loss.numel() / valid.numel()
            loss = (loss * valid.float()).sum() / (valid.float().sum() * elem_ratio + eps)
        if self.use_for_loss:
            loss = 1.0 - loss
        return loss


class AlignedL2(nn.Module):
    """L2 error after spatial and colour alignment of the prediction to the ground truth (reference :166-191), over the valid
    pixels of the whole batch; the two masked sums come from one fused launch (`dbsr_mse_per_image` with the mask)."""
    def __init__(self, alignment_net, sr_factor=4, boundary_ignore=None):
        super().__init__()
        self.sca = sca_utils.SpatialColorAlignment(alignment_net, sr_factor)
        self.boundary_ignore = boundary_ignore

    def forward(self, pred, gt, burst_input):
        aligned, valid = self.sca(pred, gt, burst_input)
        sums = ops.mse_per_image(aligned.contiguous(), gt.contiguous(), crop=self.boundary_ignore or 0, valid=valid, raw=True)
        return sums[:, 0].sum() / (sums[:, 1].sum() + 1e-12)
