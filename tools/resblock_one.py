"""Time the fused HR residual block against the two-launch path at bench shape (profiling aid):
    python tools/resblock_one.py [B] [S] [fused|two]      (maps [B, 8S, 8S, 32])"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200 import ops
from deep_rawburst_sr_b200.engine import pack_tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 48
mode = sys.argv[3] if len(sys.argv) > 3 else 'fused'
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
H = W = 8 * S
x = ops.Act(torch.randn(B, H, W, 32, generator=g).to(dev).bfloat16())
t = ops.Act.empty(B, H, W, 32, torch.bfloat16, dev)
y = ops.Act.empty(B, H, W, 32, torch.bfloat16, dev)
w1 = pack_tc((torch.randn(32, 32, 3, 3, generator=g) / 17).to(dev))
w2 = pack_tc((torch.randn(32, 32, 3, 3, generator=g) / 17).to(dev))
b1, b2 = torch.randn(32, generator=g).to(dev), torch.randn(32, generator=g).to(dev)


def run():
    if mode == 'fused':
        ops.resblock32_tc(x, y, w1, b1, w2, b2)
    else:
        ops.conv2d(x, w1, b1, t, 3, 1, 1, ops.ACT_RELU, None, tensor_core=True)
        ops.conv2d(t, w2, b2, y, 3, 1, 1, ops.ACT_RELU, x, tensor_core=True)


evs = []
for i in range(6):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); e.record(); evs.append((a, e))
torch.cuda.synchronize()
ms = min(a.elapsed_time(e) for a, e in evs[1:])
fl = 2 * 2 * B * H * W * 32 * 32 * 9
print(f'resblock_one {mode} B={B} {H}x{W}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  {2 * B * H * W * 64 / ms / 1e6:.0f} GB/s (x in + y out)')
