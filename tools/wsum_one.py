"""Run the fused warp + softmax + weighted-sum kernel at bench shape (profiling aid): python tools/wsum_one.py [B]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, C, H, W = 14, 512, 48, 48
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
feat = ops.Act(torch.rand(B * N, H, W, C, generator=g).to(dev).bfloat16())
logits = ops.Act(torch.randn(B * N, H, W, C, generator=g).to(dev).bfloat16())
offs = ((torch.rand(B * (N - 1), 2, H, W, generator=g) * 2 - 1) * 0.8).to(dev)
fused = ops.Act.empty(B, H, W, C, torch.bfloat16, dev)
evs = []
for i in range(5):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.softmax_wsum(feat, logits, fused, N, offsets=offs); e.record(); evs.append((a, e))
torch.cuda.synchronize()
ms = min(a.elapsed_time(e) for a, e in evs[1:])
nbytes = B * ((2 * N * C * 2 + C * 2) * H * W + (N - 1) * 2 * 4 * H * W)
print(f'wsum_one B={B}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.1f} GB/s algorithmic')
