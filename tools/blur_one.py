"""blur3x3 of the decoder's high-resolution map alone: python tools/blur_one.py [images] (384 x 384 x 32 bf16), CUDA-graph
of 20 launches on rotating buffers larger than L2."""
import sys
import torch
from deep_rawburst_sr_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda', 0)
nbuf = max(2, int(300e6 // (n * 384 * 384 * 32 * 2)) + 1)
xs = [ops.Act(torch.randn((n, 384, 384, 32), device=dev).bfloat16()) for _ in range(nbuf)]
ys = [ops.Act(torch.empty((n, 384, 384, 32), dtype=torch.bfloat16, device=dev)) for _ in range(nbuf)]
g1 = torch.tensor([0.25, 0.5, 0.25])
k9 = (g1[:, None] * g1[None, :]).reshape(-1).tolist()
reps = 20


def run():
    for i in range(reps):
        ops.blur3x3(xs[i % nbuf], ys[i % nbuf], k9)


side = torch.cuda.Stream()
with torch.cuda.stream(side):
    run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / (5 * reps) * 1e3
nbytes = 2 * n * 384 * 384 * 32 * 2
print('blur_one n=%d: %.1f us  %.0f GB/s (read + write)' % (n, us, nbytes / us * 1e-3))
