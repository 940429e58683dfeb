"""Spatial + colour alignment of a prediction to the ground truth (BurstSR metrics), with the reference's interface
(models/loss/spatial_color_alignment.py:23-108): `match_colors(im_ref, im_q, im_test, ksz, gauss_kernel)` and
`SpatialColorAlignment(alignment_net, sr_factor=4).forward(pred, gt, burst_input) -> (pred_warped_m, valid)`.

SURVEY.md 8(f) rank 1: the metric runs PWC-Net AGAIN at the output resolution (640^2 -> level-2 maps 160^2) plus two
warps -- those are the sm_100a kernels of the hot path (`PWCNet` -> DBSREngine's extractor / cost volume / decoder /
refiner kernels, `warp` -> dbsr_warp).  The rest is a few hundred kB of per-image glue (7x7 Gaussian on the 80x80 LR
frames, a 3x3 least-squares colour matrix, thresholds, x8 resizes of a mask) and stays in torch ops ON THE DEVICE,
written so that no TF32 path is involved (explicit tap sums and broadcasts instead of cudnn conv / matmul).  CPU tensors
are refused like everywhere else in this package.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ..layers import warp as lispr_warp
from ..layers.filtering import apply_kernel, get_gaussian_kernel


def _color_apply(im, c_mat):
    """im [B, 3, H, W], c_mat [B, 3, 3]: out[b, j] = sum_i im[b, i] * c_mat[b, i, j]  (the reference's
    matmul(im_re.permute(0, 2, 1), c_mat), spatial_color_alignment.py:45,66) as an exact-fp32 broadcast."""
    return (im.unsqueeze(2) * c_mat.unsqueeze(-1).unsqueeze(-1)).sum(dim=1)


_COLOR_ERR_THRESHOLD = 20      # on the 0..255 scale (:48-50)
_TRIM = 5                      # pixels dropped on every side after the 7x7 blur (:27-31)


def _fit_color_matrix(src, dst):
    """Per image, the 3x3 matrix X minimising || src^T X - dst^T || over the pixels (src, dst: [B, 3, h, w]) -- what the
    reference obtains from `torch.lstsq(dst.t(), src.t())` (:40-42; removed from torch 2.x).  Solved for the whole batch as
    fp64 normal equations of the [P, 3] systems: exact enough, and no host round trip."""
    A = src.flatten(2).transpose(1, 2).double()
    Bm = dst.flatten(2).transpose(1, 2).double()
    At = A.transpose(1, 2)
    return torch.linalg.solve(At @ A, At @ Bm).float()


def match_colors(im_ref, im_q, im_test, ksz, gauss_kernel):
    """Estimates a colour transformation matrix between im_ref and im_q and applies it to im_test; also returns the mask of
    pixels whose blurred colours agree after the transformation (spatial_color_alignment.py:23-69)."""
    ops.require_device(im_ref)
    kernel = gauss_kernel.to(im_ref.device)
    t = _TRIM
    ref_blur = apply_kernel(im_ref, ksz, kernel)[:, :, t:-t, t:-t].contiguous()
    qry_blur = apply_kernel(im_q, ksz, kernel)[:, :, t:-t, t:-t].contiguous()
    c_mat = _fit_color_matrix(qry_blur, ref_blur)
    # colour error of the fit, as a mask at the resolution of im_test
    residual = (_color_apply(qry_blur, c_mat) - ref_blur) * 255.0
    agree = residual.norm(dim=1) < _COLOR_ERR_THRESHOLD
    border = (im_q.shape[-1] - agree.shape[-1]) // 2
    agree = F.pad(agree, [border] * 4)
    scale = im_test.shape[-1] / agree.shape[-1]
    valid = F.interpolate(agree.unsqueeze(1).float(), scale_factor=scale, mode='bilinear') > 0.9
    return _color_apply(im_test, c_mat), valid


class SpatialColorAlignment(nn.Module):
    def __init__(self, alignment_net, sr_factor=4):
        super().__init__()
        self.sr_factor = sr_factor
        self.alignment_net = alignment_net
        self.gauss_kernel, self.ksz = get_gaussian_kernel(sd=1.5)
        # extension: normalise every image of a batch by ITS OWN maximum before the alignment net, so that a batch is
        # aligned exactly as the reference's batch-1 evaluation loop aligns its images one by one (False: the reference
        # expression, one maximum over the whole batch, spatial_color_alignment.py:88)
        self.per_image_norm = False

    def to(self, device):
        """ Move the network to device (reference signature: returns None, spatial_color_alignment.py:80-87) """
        self.alignment_net.to(device)
        self.gauss_kernel = self.gauss_kernel.to(device)

    @ops.tensor_device_guard
    @torch.no_grad()
    def forward(self, pred, gt, burst_input):
        """METRIC-ONLY (inference path): the whole alignment runs without autograd.  The reference disables gradients for the
        flow estimate only and keeps warp / colour matching differentiable w.r.t. `pred` because `AlignedL2` is also its
        BurstSR TRAINING loss (spatial_color_alignment.py:88-90); the backward pass is out of this tier's scope, so a
        prediction that requires grad is refused instead of silently returning a tensor without a graph."""
        if torch.is_tensor(pred) and pred.requires_grad:
            raise NotImplementedError('SpatialColorAlignment / AlignedL2 of deep_rawburst_sr_b200 are metric-only (no autograd '
                                      'through the sm_100a kernels): detach the prediction, or use the reference for training')
        ops.require_device(pred)
        # flow between the prediction and the ground truth: PWC-Net at the output resolution on the sm_100a kernels
        if getattr(self, 'per_image_norm', False):
            flow = self.alignment_net(pred / (pred.amax(dim=(1, 2, 3), keepdim=True) + 1e-6),
                                      gt / (gt.amax(dim=(1, 2, 3), keepdim=True) + 1e-6))
        else:
            flow = self.alignment_net(pred / (pred.max() + 1e-6), gt / (gt.max() + 1e-6))
        pred_warped = lispr_warp.warp(pred, flow)
        sr_factor = self.sr_factor
        ds_factor = 1.0 / float(2.0 * sr_factor)
        flow_ds = F.interpolate(flow, scale_factor=ds_factor, mode='bilinear') * ds_factor
        burst_0 = burst_input[:, 0, [0, 1, 3]].contiguous()
        burst_0_warped = lispr_warp.warp(burst_0, flow_ds)
        frame_gt_ds = F.interpolate(gt, scale_factor=ds_factor, mode='bilinear')
        pred_warped_m, valid = match_colors(frame_gt_ds, burst_0_warped, pred_warped, self.ksz, self.gauss_kernel)
        return pred_warped_m, valid
