"""One SSIM launch chain (32 x 3 x 384 x 384, BASELINE configs[1] output size) for `ncu -k regex:ssim_tile`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200.models.loss import msssim as ms  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
gt = torch.rand(32, 3, 384, 384, generator=g).to(dev)
pred = (gt + 0.02 * torch.randn(32, 3, 384, 384, generator=g).to(dev)).clamp(0, 1)
for _ in range(4):
    s = ms.ssim(pred, gt, size_average=False)
torch.cuda.synchronize()
print('ssim per image', s[:4].tolist())
