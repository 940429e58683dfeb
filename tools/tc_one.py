"""Run one tcgen05 conv shape a few times (profiling aid for ncu): python tools/tc_one.py cin cout k n h w [residual] [reps]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_rawburst_sr_b200 import ops
from deep_rawburst_sr_b200.engine import pack_tc

cin, cout, k, n, h, w = [int(v) for v in sys.argv[1:7]]
use_res = len(sys.argv) > 7 and sys.argv[7] == '1'
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
pitch = max(8, (cin + 7) // 8 * 8)
x = ops.Act((torch.randn(n, h, w, pitch, generator=g) * 0.5).to(dev).bfloat16()).slice(0, cin)
wt = pack_tc((torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev))
b = torch.randn(cout, generator=g).to(dev)
y = ops.Act.empty(n, h, w, cout, torch.bfloat16, dev)
r = ops.Act((torch.randn(n, h, w, cout, generator=g)).to(dev).bfloat16()) if use_res else None
evs = []
for i in range(reps):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.conv2d(x, wt, b, y, k, 1, 1, ops.ACT_RELU, r, 0, tensor_core=True)
    e.record()
    evs.append((a, e))
torch.cuda.synchronize()
ms = min(a.elapsed_time(e) for a, e in evs[1:])
fl = 2.0 * n * h * w * cout * cin * k * k
print(f'tc_one cin={cin} cout={cout} k={k} n={n} {h}x{w} res={use_res}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s')
