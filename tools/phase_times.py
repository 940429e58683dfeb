"""Where a forward spends its time: CUDA-graph replays of the alignment chain (PWC-Net), the encoder stack and the
fusion + decoder tail, each captured alone, against the whole forward (two-stream fork / join).
usage: python tools/phase_times.py [batch ...]"""
import sys
import torch
from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic


def timed(graph, reps=50):
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def capture(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    n0 = eng.launches
    with torch.cuda.graph(g):
        fn()
    return g, eng.launches - n0


if __name__ == '__main__':
    batches = [int(a) for a in sys.argv[1:]] or [1, 2, 32]
    torch.manual_seed(0)
    net = dbsrnet_default_synthetic().cuda().eval().set_precision('bf16')
    eng = net.engine(torch.device('cuda', 0))
    N, H, W = 14, 48, 48
    for B in batches:
        burst = torch.rand(B, N, 4, H, W, device='cuda')
        ws = eng.workspace((B, N, H, W))
        enc_in = eng._buf(ws, 'enc_in', B * N, H, W, 8, eng.act_dtype)
        offsets = torch.zeros((B * (N - 1), 2, H, W), device='cuda')
        pred = torch.empty((B, 3, 8 * H, 8 * W), device='cuda')
        state = {}

        def f_pwc():
            s2d0, pwc_in = eng.prep(ws, burst, enc_in)
            eng.pwc_burst(ws, pwc_in, B, N, H, W, offsets, s2d0)

        def f_enc():
            eng.prep(ws, burst, enc_in)
            state['feat'] = eng.encode(ws, enc_in)
            eng.project(ws, state['feat'])

        def f_tail():
            fused = eng.merge(ws, state['feat'], offsets, B, N, None, aligned=False, projected=True)
            eng.decode(ws, fused, pred)

        def f_full():
            eng.forward(burst, out={'pred': pred})

        with torch.no_grad():
            res = {}
            for name, fn in (('pwc', f_pwc), ('encoder', f_enc), ('tail', f_tail), ('full', f_full)):
                g, nl = capture(fn)
                res[name] = (timed(g), nl)
        print('B=%d  ' % B + '  '.join('%s %.0f us (%d launches)' % (k, v[0], v[1]) for k, v in res.items()), flush=True)
