"""CPU oracle of the evaluation metrics -- TEST INFRASTRUCTURE ONLY.

Restates, with explicit shifted-slice arithmetic in fp32 (no library convolution / pooling), the reference's
`gaussian` / `create_window` / `ssim` / `msssim` (models/loss/msssim.py:10-104), `msssim.SSIM.forward` (:119-130),
`image_quality_v2.SSIM.forward` (models/loss/image_quality_v2.py:111-136) and `PSNR` (:69-101).  Same rules as
`dbsr_oracle`: only tests / smoke / the CPU baseline legs may import it.

Pin: `oracle/make_golden_metrics.py` runs the reference's own modules (`lpips`, imported by image_quality_v2.py:21 and not
on this path, stubbed) and commits `tests/golden/metrics_*.npz`; `tests/test_oracle.py` checks this file against them.
"""
from __future__ import annotations

import math

import torch


def make_image_pair(seed: int, n: int, c: int, h: int, w: int, noise: float = 0.05, scale: float = 1.0, offset: float = 0.0):
    """Seeded (pred, gt): a low-pass random image and a perturbed copy, so that local variances span smooth to textured."""
    g = torch.Generator().manual_seed(7000 + seed)
    base = torch.rand(n, c, h // 4 + 2, w // 4 + 2, generator=g)
    ys = torch.linspace(0.5, h // 4 + 0.5, h)
    xs = torch.linspace(0.5, w // 4 + 0.5, w)
    y0, x0 = ys.floor().long(), xs.floor().long()
    fy, fx = (ys - y0).view(-1, 1), (xs - x0).view(1, -1)
    gt = ((1 - fy) * (1 - fx) * base[..., y0, :][..., x0] + (1 - fy) * fx * base[..., y0, :][..., x0 + 1] +
          fy * (1 - fx) * base[..., y0 + 1, :][..., x0] + fy * fx * base[..., y0 + 1, :][..., x0 + 1])
    gt = (gt + 0.1 * torch.rand(n, c, h, w, generator=g)).clamp(0, 1)
    pred = (gt + noise * torch.randn(n, c, h, w, generator=g)).clamp(0, 1)
    return (pred * scale + offset).contiguous(), (gt * scale + offset).contiguous()


def gaussian(window_size: int, sigma: float = 1.5):
    """msssim.py:10-12 (python doubles -> fp32 tensor -> normalised in fp32)"""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=torch.float32)
    return g / g.sum()


def window2d(window_size: int):
    """msssim.py:15-19: fp32 outer product of the 1-D window"""
    g = gaussian(window_size)
    return g.view(-1, 1) * g.view(1, -1)


def _filter_valid(x, w2):
    """F.conv2d(x, window, padding=0, groups=C) (msssim.py:44-45): the same 2-D window on every channel, no padding"""
    k = w2.shape[0]
    oh, ow = x.shape[-2] - k + 1, x.shape[-1] - k + 1
    out = torch.zeros(x.shape[:-2] + (oh, ow), dtype=torch.float32)
    for i in range(k):
        for j in range(k):
            out = out + w2[i, j] * x[..., i:i + oh, j:j + ow]
    return out


def value_range(img1):
    """msssim.py:24-35"""
    max_val = 255 if float(img1.max()) > 128 else 1
    min_val = -1 if float(img1.min()) < -0.5 else 0
    return max_val - min_val


def ssim_map(img1, img2, window_size: int = 11, val_range=None, fixed_window: bool = False):
    """msssim.py:22-63 -> (ssim_map, v1 / v2).  fixed_window: the caller hands a window_size window in (the SSIM class)."""
    L = value_range(img1) if val_range is None else val_range
    k = window_size if fixed_window else min(window_size, img1.shape[-2], img1.shape[-1])
    w2 = window2d(k)
    mu1, mu2 = _filter_valid(img1, w2), _filter_valid(img2, w2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = _filter_valid(img1 * img1, w2) - mu1_sq
    sigma2_sq = _filter_valid(img2 * img2, w2) - mu2_sq
    sigma12 = _filter_valid(img1 * img2, w2) - mu1_mu2
    C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    v1 = 2.0 * sigma12 + C2
    v2 = sigma1_sq + sigma2_sq + C2
    return ((2 * mu1_mu2 + C1) * v1) / ((mu1_sq + mu2_sq + C1) * v2), v1 / v2


def ssim(img1, img2, window_size: int = 11, size_average: bool = True, full: bool = False, val_range=None, spatial_out: bool = False):
    """msssim.py:22-74"""
    m, csm = ssim_map(img1, img2, window_size, val_range)
    cs = csm.mean()
    ret = m if spatial_out else (m.mean() if size_average else m.mean(1).mean(1).mean(1))
    return (ret, cs) if full else ret


def avg_pool2(x):
    """F.avg_pool2d(x, (2, 2)) (msssim.py:88-89): floor output size"""
    oh, ow = x.shape[-2] // 2, x.shape[-1] // 2
    x = x[..., :2 * oh, :2 * ow]
    return (((x[..., 0::2, 0::2] + x[..., 0::2, 1::2]) + x[..., 1::2, 0::2]) + x[..., 1::2, 1::2]) * 0.25


def msssim(img1, img2, window_size: int = 11, val_range=None, normalize: bool = False):
    """msssim.py:77-104 (size_average=True)"""
    weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float32)
    sims, css = [], []
    for _ in range(5):
        s, c = ssim(img1, img2, window_size, True, True, val_range)
        sims.append(s)
        css.append(c)
        img1, img2 = avg_pool2(img1), avg_pool2(img2)
    sims, css = torch.stack(sims), torch.stack(css)
    if normalize:
        sims, css = (sims + 1) / 2, (css + 1) / 2
    return torch.prod((css ** weights)[:-1] * (sims ** weights)[-1])


def ssim_metric(pred, gt, boundary_ignore=None, use_for_loss: bool = True, valid=None):
    """image_quality_v2.py:104-136 (`SSIM`): crop, 11-tap window, value range always data-derived (msssim.py:129-130)"""
    if boundary_ignore is not None:
        b = boundary_ignore
        pred, gt = pred[..., b:-b, b:-b], gt[..., b:-b, b:-b]
        if valid is not None:
            valid = valid[..., b:-b, b:-b]
    if pred.dim() == 3:
        pred, gt = pred.unsqueeze(0), gt.unsqueeze(0)
    m, _ = ssim_map(pred, gt, 11, None, fixed_window=True)
    if valid is not None:
        valid = valid[..., 5:-5, 5:-5]
        ratio = m.numel() / valid.numel()
        loss = (m * valid.float()).sum() / (valid.float().sum() * ratio + 1e-12)
    else:
        loss = m.mean()
    return 1.0 - loss if use_for_loss else loss


def psnr_per_image(pred, gt, boundary_ignore=None, max_value: float = 1.0):
    """image_quality_v2.py:47-66 (l2, valid=None) and :75-86, one value per image"""
    if boundary_ignore is not None:
        b = boundary_ignore
        pred, gt = pred[..., b:-b, b:-b], gt[..., b:-b, b:-b]
    mse = ((pred - gt) ** 2).flatten(1).mean(1)
    return 20 * math.log10(max_value) - 10.0 * mse.log10()


def psnr(pred, gt, boundary_ignore=None, max_value: float = 1.0):
    """image_quality_v2.py:88-101: per image, inf / nan dropped, averaged"""
    p = psnr_per_image(pred, gt, boundary_ignore, max_value)
    p = p[torch.isfinite(p)]
    return p.mean() if p.numel() else torch.zeros(())
