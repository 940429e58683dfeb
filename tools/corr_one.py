"""Run the tiled cost-volume kernel at one pyramid-level shape (profiling aid): python tools/corr_one.py [C h pairs warp]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200 import ops
from deep_rawburst_sr_b200.ops import ACT_LRELU, Act
a = [int(v) for v in sys.argv[1:]]
C, h, P, warp = (a + [32, 32, 104, 1][len(a):])[:4]
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16


def padded(t, dtype):
    n, hh, ww, c = t.shape
    b = Act.empty(n, hh, ww, (c + 7) // 8 * 8, dtype, dev, zero=True)
    b.buf[..., :c] = t.to(dev).to(dtype)
    return b.slice(0, c)


f1 = padded(torch.randn(P, h, h, C, generator=g), dt)
f2 = padded(torch.randn(P, h, h, C, generator=g), dt)
vol = Act.empty(P, h, h, 88, dt, dev, zero=True).slice(0, 81)
flow = padded((torch.rand(P, h, h, 2, generator=g) * 2 - 1) * 0.5, torch.float32) if warp else None
evs = []
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for i in range(5):
    flush.zero_()      # ~100 us of queued GPU work: the host runs ahead, the events bracket the kernel and not the Python launch path
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.corr81(f1, f2, vol, P, 0, flow=flow, flow_scale=1.0 if warp else 0.0, act=ACT_LRELU); e.record(); evs.append((s, e))
torch.cuda.synchronize()
ms = min(s.elapsed_time(e) for s, e in evs[1:])
nbytes = P * (2 * C + 81) * h * h * 2
print(f'corr_one C={C} h={h} pairs={P} warp={warp}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.1f} GB/s algorithmic')
