"""CPU oracle of the BurstSR spatial + colour alignment metric path -- TEST INFRASTRUCTURE ONLY.

Restates `SpatialColorAlignment.forward` / `match_colors` (reference models/loss/spatial_color_alignment.py:23-108),
`get_gaussian_kernel` / `apply_kernel` (models/layers/filtering.py:43-62) and `AlignedL2.forward`
(models/loss/image_quality_v2.py:166-191) with explicit coordinate arithmetic on torch CPU tensors, on top of the
PWC-Net / warp / resize restatements of `dbsr_oracle`.  Same rules as `dbsr_oracle`: only tests / smoke / the CPU
baseline legs may import it.

Pin: `oracle/make_golden_sca.py` runs the reference's own modules (with `torch.lstsq`, removed from torch 2.x, shimmed
by `torch.linalg.lstsq` -- the only change) and commits `tests/golden/sca_*.npz`; `tests/test_oracle.py` checks this
file against them.
"""
from __future__ import annotations

import math

import torch

from . import dbsr_oracle as O


def gaussian_kernel(sd: float, ksz=None):
    """filtering.py:43-52 (gauss_2d with density=True, normalised to sum 1); ksz = int(4 sd + 1) must be odd."""
    if ksz is None:
        ksz = int(4 * sd + 1)
    assert ksz % 2 == 1
    k = torch.arange(-(ksz - 1) / 2, (ksz + 1) / 2)
    g = torch.exp(-1.0 / (2 * sd ** 2) * k ** 2) / (math.sqrt(2 * math.pi) * sd)
    K = g.view(1, -1) * g.view(-1, 1)
    return K / K.sum(), ksz


def apply_kernel(im, ksz: int, K):
    """filtering.py:55-62: per-channel correlation with K after REFLECT padding by ksz // 2 (explicit index arithmetic)."""
    H, W = im.shape[-2:]
    r = ksz // 2

    def refl(i, n):
        i = i.abs()
        return torch.where(i > n - 1, 2 * (n - 1) - i, i)

    ys, xs = torch.arange(H), torch.arange(W)
    out = torch.zeros_like(im)
    for dy in range(ksz):
        yy = refl(ys + dy - r, H)
        for dx in range(ksz):
            xx = refl(xs + dx - r, W)
            out = out + K[dy, dx] * im[..., yy, :][..., xx]
    return out


def resize_scale(x, factor: float):
    """F.interpolate(x, scale_factor=factor, mode='bilinear') for sizes the factor divides evenly (then the coordinate
    scale 1/factor equals in/out and dbsr_oracle.resize_bilinear applies): spatial_color_alignment.py:61,93,97."""
    H, W = x.shape[-2:]
    Ho, Wo = int(math.floor(H * factor)), int(math.floor(W * factor))
    assert abs(Ho / H - factor) < 1e-12 and abs(Wo / W - factor) < 1e-12, 'oracle covers evenly dividing factors only'
    return O.resize_bilinear(x, Ho, Wo)


def match_colors(im_ref, im_q, im_test, ksz, K):
    """spatial_color_alignment.py:23-69."""
    bi = 5
    ref_mean = apply_kernel(im_ref, ksz, K)[:, :, bi:-bi, bi:-bi]
    q_mean = apply_kernel(im_q, ksz, K)[:, :, bi:-bi, bi:-bi]
    B = im_ref.shape[0]
    ref_re = ref_mean.reshape(B, 3, -1)
    q_re = q_mean.reshape(B, 3, -1)
    c_all = []
    for ir, iq in zip(ref_re, q_re):
        # torch.lstsq(ir.t(), iq.t()).solution[:3]  =  argmin_X || iq.t() X - ir.t() ||   (:40-42)
        c_all.append(torch.linalg.lstsq(iq.t().double(), ir.t().double()).solution.float())
    c_mat = torch.stack(c_all, 0)                                        # [B, 3, 3]
    q_conv = torch.matmul(q_re.permute(0, 2, 1), c_mat).permute(0, 2, 1).reshape(q_mean.shape)
    err = ((q_conv - ref_mean) * 255.0).norm(dim=1)
    valid = err < 20                                                     # :50-53
    pad = (im_q.shape[-1] - valid.shape[-1]) // 2
    valid = torch.nn.functional.pad(valid, [pad, pad, pad, pad])
    up = im_test.shape[-1] / valid.shape[-1]
    valid = resize_scale(valid.unsqueeze(1).float(), up) > 0.9           # :59-62
    t_re = im_test.reshape(B, 3, -1)
    t_conv = torch.matmul(t_re.permute(0, 2, 1), c_mat).permute(0, 2, 1).reshape(im_test.shape)
    return t_conv, valid, c_mat


def spatial_color_alignment(pred, gt, burst_input, pwc_sd, sr_factor: int = 4, pre: str = 'net.'):
    """SpatialColorAlignment.forward, spatial_color_alignment.py:85-108.  pwc_sd: PWCNet state_dict (keys 'net.*')."""
    K, ksz = gaussian_kernel(1.5)
    flow = O.pwcnet_forward(pred / (pred.max() + 1e-6), gt / (gt.max() + 1e-6), pwc_sd, pre=pre)
    pred_warped = O.warp(pred, flow)
    ds = 1.0 / float(2.0 * sr_factor)
    flow_ds = resize_scale(flow, ds) * ds
    burst_0 = burst_input[:, 0, [0, 1, 3]].contiguous()
    burst_0_warped = O.warp(burst_0, flow_ds)
    gt_ds = resize_scale(gt, ds)
    pred_m, valid, c_mat = match_colors(gt_ds, burst_0_warped, pred_warped, ksz, K)
    return pred_m, valid, {'flow': flow, 'c_mat': c_mat, 'pred_warped': pred_warped}


def aligned_l2(pred, gt, burst_input, pwc_sd, sr_factor: int = 4, boundary_ignore=None, pre: str = 'net.'):
    """AlignedL2.forward, image_quality_v2.py:173-191."""
    pred_m, valid, _ = spatial_color_alignment(pred, gt, burst_input, pwc_sd, sr_factor, pre)
    if boundary_ignore is not None:
        b = boundary_ignore
        pred_m, gt, valid = pred_m[..., b:-b, b:-b], gt[..., b:-b, b:-b], valid[..., b:-b, b:-b]
    mse = (pred_m - gt) ** 2
    ratio = mse.numel() / valid.numel()
    return (mse * valid.float()).sum() / (valid.float().sum() * ratio + 1e-12)


def make_sca_inputs(seed: int, B: int, S: int, sr_factor: int = 4):
    """Seeded, structured inputs: a smooth random scene; `pred` = the scene shifted by one HR pixel, colour-transformed and
    noised; `burst_input[:, 0]` = the scene downsampled by 2*sr_factor in another colour space, packed as RGGB."""
    g = torch.Generator().manual_seed(4000 + seed)
    f = 2 * sr_factor
    assert S % f == 0
    coarse = torch.rand(B, 3, S // 16 + 4, S // 16 + 4, generator=g)
    canvas = torch.nn.functional.interpolate(coarse, size=(S + 32, S + 32), mode='bicubic', align_corners=False).clamp(0.02, 0.98)
    gt = canvas[..., 16:16 + S, 16:16 + S].contiguous()
    M = torch.eye(3) + 0.08 * torch.randn(3, 3, generator=g)
    shifted = canvas[..., 17:17 + S, 15:15 + S]
    pred = (torch.einsum('ij,bjhw->bihw', M, shifted) + 0.01 * torch.randn(B, 3, S, S, generator=g)).clamp(0.0, 1.0).contiguous()
    lr = torch.nn.functional.avg_pool2d(gt, f)
    Mq = torch.eye(3) + 0.1 * torch.randn(3, 3, generator=g)
    lr_q = torch.einsum('ij,bjhw->bihw', Mq, lr).clamp(0.0, 1.0)
    frame0 = torch.stack([lr_q[:, 0], lr_q[:, 1], lr_q[:, 1], lr_q[:, 2]], 1)
    burst = torch.stack([frame0, torch.rand(B, 4, S // f, S // f, generator=g)], 1).contiguous()
    # corrupt a patch of frame 0 so that some pixels fail the colour-error threshold (valid == False somewhere)
    burst[:, 0, :, 8:12, 8:13] = 1.0 - burst[:, 0, :, 8:12, 8:13]
    return pred, gt, burst


def pwc_state_dict(seed: int = 0, gain: float = 1.0):
    """PWCNet module state_dict ('net.*' keys) carved out of dbsr_oracle.make_state_dict."""
    pre = 'encoder.alignment_net.'
    return {k[len(pre):]: v for k, v in O.make_state_dict(seed, pwc_gain=gain).items() if k.startswith(pre)}
