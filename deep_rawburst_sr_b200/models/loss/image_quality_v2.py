"""Image-quality metrics of the DBSR path with the reference's interface (models/loss/image_quality_v2.py):
`PixelWiseError` (:24-66), `PSNR` (:69-101) and `AlignedL2` (:166-191, the BurstSR loss / metric built on
`SpatialColorAlignment`) and `SSIM` (:104-136, on the fused `dbsr_ssim` kernel of `msssim.py`).  On CUDA fp32 batches
without a `valid` mask `PSNR` takes its per-image MSE from one `dbsr_mse_per_image` launch (boundary_ignore as index
arithmetic) instead of a Python loop of sliced reductions.  LPIPS depends on a package outside this path (`lpips`, a
pretrained AlexNet) and is not provided."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from . import msssim
from . import spatial_color_alignment as sca_utils


class PixelWiseError(nn.Module):
    """ Computes pixel-wise error using the specified metric. Optionally boundary pixels are ignored during error
        calculation """
    def __init__(self, metric='l1', boundary_ignore=None):
        super().__init__()
        self.boundary_ignore = boundary_ignore
        if metric == 'l1':
            self.loss_fn = F.l1_loss
        elif metric == 'l2':
            self.loss_fn = F.mse_loss
        elif metric == 'l2_sqrt':
            self.loss_fn = lambda pred, gt: (((pred - gt) ** 2).sum(dim=-3)).sqrt().mean()
        elif metric == 'charbonnier':
            self.loss_fn = lambda pred, gt: ((pred - gt) ** 2 + 1e-3 ** 2).sqrt().mean()
        else:
            raise Exception

    def forward(self, pred, gt, valid=None):
        if self.boundary_ignore is not None:
            b = self.boundary_ignore
            pred = pred[..., b:-b, b:-b]
            gt = gt[..., b:-b, b:-b]
            if valid is not None:
                valid = valid[..., b:-b, b:-b]
        if valid is None:
            return self.loss_fn(pred, gt)
        err = self.loss_fn(pred, gt, reduction='none')
        eps = 1e-12
        elem_ratio = err.numel() / valid.numel()
        return (err * valid.float()).sum() / (valid.float().sum() * elem_ratio + eps)


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
    """ Computes L2 error after performing spatial and color alignment of the input image to GT"""
    def __init__(self, alignment_net, sr_factor=4, boundary_ignore=None):
        super().__init__()
        self.sca = sca_utils.SpatialColorAlignment(alignment_net, sr_factor)
        self.boundary_ignore = boundary_ignore

    def forward(self, pred, gt, burst_input):
        pred_warped_m, valid = self.sca(pred, gt, burst_input)
        if self.boundary_ignore is not None:
            b = self.boundary_ignore
            pred_warped_m = pred_warped_m[..., b:-b, b:-b]
            gt = gt[..., b:-b, b:-b]
            valid = valid[..., b:-b, b:-b]
        mse = F.mse_loss(pred_warped_m, gt, reduction='none')
        eps = 1e-12
        elem_ratio = mse.numel() / valid.numel()
        return (mse * valid.float()).sum() / (valid.float().sum() * elem_ratio + eps)
