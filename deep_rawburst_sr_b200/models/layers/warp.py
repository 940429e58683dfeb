"""`warp(feat, flow)` with the reference's signature (models/layers/warp.py:19-46), backed by `dbsr_warp`."""
import torch

from ... import ops


def warp(feat, flow, mode='bilinear', padding_mode='zeros'):
    """feat [B, C, H, W], flow [B, 2, H, W] (x, y) -> feat sampled at (x + flow_x, y + flow_y), zeros outside."""
    if mode != 'bilinear' or padding_mode != 'zeros':
        raise NotImplementedError('deep_rawburst_sr_b200.warp supports bilinear / zeros only')
    ops.require_device(feat)
    n, c, h, w = feat.shape
    cp = (c + 3) // 4 * 4
    src = ops.Act.empty(n, h, w, cp, torch.float32, feat.device, zero=cp != c)
    src.slice(0, c).from_nchw(feat.contiguous().float())
    dst = ops.Act.empty(n, h, w, cp, torch.float32, feat.device)
    ops.warp(src, flow.contiguous().float(), dst, frames=0)
    return dst.slice(0, c).to_nchw()
