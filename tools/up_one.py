"""Time the decoder's upsampling conv (1x1 64 -> 2048 + ReLU + PixelShuffle(8) folded into the store) and the 3x3 blur at
bench shape (profiling aid): python tools/up_one.py [B] [S]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200 import ops
from deep_rawburst_sr_b200.engine import pack_tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 48
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
x = ops.Act(torch.randn(B, S, S, 64, generator=g).to(dev).bfloat16())
w = pack_tc((torch.randn(2048, 64, 1, 1, generator=g) / 8).to(dev), shuffle_r=8)
y = ops.Act.empty(B, 8 * S, 8 * S, 32, torch.bfloat16, dev)
z = ops.Act.empty(B, 8 * S, 8 * S, 32, torch.bfloat16, dev)
k = torch.tensor([1.0, 2.0, 1.0]); k = (k[:, None] * k[None, :] / 16).reshape(-1).tolist()
for name, fn in (('upsample', lambda: ops.conv2d(x, w, None, y, 1, 1, 1, ops.ACT_RELU, None, 8, tensor_core=True)),
                 ('blur3x3', lambda: ops.blur3x3(y, z, k))):
    evs = []
    for i in range(6):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); evs.append((a, e))
    torch.cuda.synchronize()
    ms = min(a.elapsed_time(e) for a, e in evs[1:])
    out_mb = B * 64 * S * S * 32 * 2 / 1e6
    print(f'up_one {name} B={B} {S}x{S}: {ms * 1e3:.1f} us  out {out_mb:.0f} MB -> {out_mb / ms / 1e3:.2f} TB/s written')
