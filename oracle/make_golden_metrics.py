"""Golden vectors of the evaluation metrics from the UNMODIFIED reference modules (build container only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_metrics.py

TEST INFRASTRUCTURE ONLY.  Imports models/loss/msssim.py and models/loss/image_quality_v2.py from /root/reference; the only
shim is a stub `lpips` module (imported at image_quality_v2.py:21, not installed, not on this path).  Inputs are the seeded
pairs of `metrics_oracle.make_image_pair`; only the metric values and a corner of each SSIM map are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path.insert(0, ROOT)

from oracle import metrics_oracle as M  # noqa: E402

# name, seed, n, c, h, w, noise, scale, offset, boundary_ignore
CASES = [
    ('metrics_rgb_b2_176', 0, 2, 3, 176, 176, 0.05, 1.0, 0.0, 8),
    ('metrics_ragged_b3_97x131', 1, 3, 3, 97, 131, 0.10, 1.0, 0.0, 4),
    ('metrics_gray_255_b1_200', 2, 1, 1, 200, 192, 0.02, 255.0, 0.0, 16),
    ('metrics_signed_b2_180', 3, 2, 3, 180, 184, 0.05, 2.0, -1.0, 6),
]


def main():
    sys.dont_write_bytecode = True
    sys.modules.setdefault('lpips', types.ModuleType('lpips'))
    sys.path.insert(0, REF)
    from models.loss import msssim as ref_ms
    from models.loss import image_quality_v2 as ref_iq
    outdir = os.path.join(ROOT, 'tests', 'golden')
    for name, seed, n, c, h, w, noise, scale, offset, bi in CASES:
        pred, gt = M.make_image_pair(seed, n, c, h, w, noise, scale, offset)
        g = torch.Generator().manual_seed(9000 + seed)
        valid = torch.rand(n, 1, h, w, generator=g) > 0.3
        with torch.no_grad():
            s_mean, cs = ref_ms.ssim(pred, gt, full=True)
            s_img = ref_ms.ssim(pred, gt, size_average=False)
            s_map = ref_ms.ssim(pred, gt, spatial_out=True)
            s_vr = ref_ms.ssim(pred, gt, val_range=1.0)
            ms = ref_ms.msssim(pred, gt)
            ms_norm = ref_ms.msssim(pred, gt, normalize=True)
            cls_mean = ref_ms.SSIM()(pred, gt)
            msc = ref_ms.MSSSIM()(pred, gt)
            iq_loss = ref_iq.SSIM(boundary_ignore=bi)(pred, gt)
            iq_val = ref_iq.SSIM(boundary_ignore=bi, use_for_loss=False)(pred, gt)
            iq_valid = ref_iq.SSIM(boundary_ignore=bi, use_for_loss=False)(pred, gt, valid)
            iq_single = ref_iq.SSIM(boundary_ignore=None, use_for_loss=False)(pred[0], gt[0])
            psnr = ref_iq.PSNR(boundary_ignore=bi, max_value=max(scale, 1.0))(pred, gt)
            psnr_each = torch.stack([ref_iq.PSNR(boundary_ignore=bi, max_value=max(scale, 1.0)).psnr(p.unsqueeze(0), q.unsqueeze(0))
                                     for p, q in zip(pred, gt)])
            small = ref_ms.ssim(pred[..., :7, :9], gt[..., :7, :9], full=True)      # real_size = min(11, 7, 9) = 7
        np.savez_compressed(
            os.path.join(outdir, name + '.npz'),
            meta=np.array([seed, n, c, h, w, bi], dtype=np.int64), gen=np.array([noise, scale, offset]),
            ssim_mean=np.float32(s_mean), cs=np.float32(cs), ssim_per_image=s_img.numpy(), ssim_map_corner=s_map[..., :24, :24].numpy(),
            ssim_map_tail=s_map[..., -16:, -16:].numpy(), ssim_val_range1=np.float32(s_vr), msssim=np.float32(ms),
            msssim_normalized=np.float32(ms_norm), ssim_class=np.float32(cls_mean), msssim_class=np.float32(msc),
            iq_ssim_loss=np.float32(iq_loss), iq_ssim=np.float32(iq_val), iq_ssim_valid=np.float32(iq_valid),
            iq_ssim_single=np.float32(iq_single), psnr=np.float32(psnr), psnr_per_image=psnr_each.numpy(),
            small_ssim=np.float32(small[0]), small_cs=np.float32(small[1]), valid_seed=np.int64(9000 + seed))
        print(name, 'ssim', float(s_mean), 'cs', float(cs), 'msssim', float(ms), 'iq', float(iq_val), float(iq_valid), 'psnr', float(psnr))


if __name__ == '__main__':
    main()
