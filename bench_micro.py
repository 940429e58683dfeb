#!/usr/bin/env python
"""Microbenchmark of the HBM-bound kernels of the DBSR path (BASELINE.json configs[3]: "PWCNet alignment + warp-fuse
microbench only: 14-frame 64-channel embeddings at 96x96 and 160x160 flow resolution"; shapes from SURVEY.md 8(d)).

    python bench_micro.py [--iters K] [--warmup W] [--bursts B] [--out file.json]

Five legs, every one through the C ABI (`ops.*` -> libdbsr_b200.so), one JSON line each (rank 0 / one GPU):

* `corr81`      cost volume (correlation.py:280-330 + the fused backwarp / LeakyReLU of pwcnet.py:161,169) at the five
                pyramid level shapes of a 96^2 and a 160^2 frame, 13*B pairs: achieved HBM GB/s = algorithmic bytes
                (f1 + f2 read once, 81-channel volume written once) / device time, against the measured HBM peak.
* `warp_fuse`   fused bilinear warp + softmax over the burst + weighted sum (warp.py:19-46, merging.py:117-124):
                C=64 at S=96,160 (the cfg-4 stress shape) and the network's real shape C=512 at S=48,80; flows
                U(-4,4) px so that out-of-image taps occur.  Bytes: (28*C*s + 13*2*4)*S^2 + C*s*S^2 per burst.
* `pwc_align`   the whole PWC-Net alignment of a burst batch (pwcnet.py:248-281) at 96^2 / 160^2: pairs/s.
* `metrics`     SSIM / MS-SSIM / PSNR of a batch of predictions through the fused metric kernels.
* `generator`   (opt-in) the synthetic burst generator: its three kernels, and `rgb2rawburst` as called (8(f) rank 4).
* `eval`        (opt-in, --legs eval) the batched SyntheticBurst scoring loop end to end from host tensors (8(f) rank 2).
* `sca`         SpatialColorAlignment.forward (models/loss/spatial_color_alignment.py:85-108) at the BurstSR evaluation
                shape (640^2 prediction / ground truth, 80^2 RAW burst): images/s.

Timing: CUDA events on the launching stream around EVERY launch, W warm-up launches, and the L2 is flushed between
timed launches (a 512 MB buffer is overwritten), so small level shapes cannot be served from the 126 MB L2.
The median launch time is reported (min next to it).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAMES = 14
EXT_CH = {2: 32, 3: 64, 4: 96, 5: 128, 6: 196}


def peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['hbm_gbs'], 'measured'
    return 6650.0, 'fallback'


class Timer:
    def __init__(self, dev, iters, warmup):
        self.iters, self.warmup = iters, warmup
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def __call__(self, fn):
        for _ in range(self.warmup):
            fn()
        evs = []
        for _ in range(self.iters):
            self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in evs)
        return t[len(t) // 2], t[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--bursts', type=int, default=8, help='bursts per launch (13 pairs each)')
    ap.add_argument('--legs', default='corr81,warp_fuse,pwc_align,sca,metrics')
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.ops import ACT_LRELU, Act
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    ops.require_device(torch.empty(1, device=dev))
    timer = Timer(dev, args.iters, args.warmup)
    peak, peak_src = peak_hbm()
    B = args.bursts
    P = B * (FRAMES - 1)
    lines = []

    def emit(d):
        d.update({'peak_gbs': peak, 'peak_source': peak_src, 'l2': 'flushed between launches (512 MB overwrite)'})
        lines.append(d)
        print(json.dumps(d), flush=True)

    legs = args.legs.split(',')
    g = torch.Generator().manual_seed(4)

    def padded(t, dt):
        """[n, h, w, c] host tensor -> device Act whose pixel pitch is rounded up to 8 channels (the engine's layout)"""
        n, h, w, c = t.shape
        a = Act.empty(n, h, w, (c + 7) // 8 * 8, dt, dev, zero=True)
        a.buf[..., :c] = t.to(dev).to(dt)
        return a.slice(0, c)
    if 'corr81' in legs:
        for S in (96, 160):
            hp = (S + 63) // 64 * 64
            for dt in (torch.bfloat16, torch.float32):
                es = 2 if dt == torch.bfloat16 else 4
                tot_b, tot_ms = 0, 0.0
                per_level = []
                for lvl in (2, 3, 4, 5, 6):
                    C, h = EXT_CH[lvl], hp >> lvl
                    f1 = padded(torch.randn(P, h, h, C, generator=g), dt)
                    f2 = padded(torch.randn(P, h, h, C, generator=g), dt)
                    vol = Act.empty(P, h, h, 88, dt, dev, zero=True).slice(0, 81)
                    flow = padded((torch.rand(P, h, h, 2, generator=g) * 2 - 1) * 0.5, torch.float32) if lvl < 6 else None
                    med, mn = timer(lambda: ops.corr81(f1, f2, vol, P, 0, flow=flow, flow_scale=1.0 if flow is not None else 0.0,
                                                       act=ACT_LRELU))
                    nbytes = P * (2 * C + 81) * h * h * es
                    per_level.append({'level': lvl, 'C': C, 'h': h, 'us': med * 1e3, 'us_min': mn * 1e3,
                                      'gbs': nbytes / med / 1e6, 'bytes': nbytes})
                    tot_b += nbytes
                    tot_ms += med
                big = per_level[0]
                emit({'leg': 'corr81', 'frame': S, 'pairs': P, 'dtype': 'bf16' if es == 2 else 'f32',
                      'algorithmic_bytes': tot_b, 'us_all_levels': tot_ms * 1e3, 'hbm_gbs_all_levels': tot_b / tot_ms / 1e6,
                      'hbm_gbs_level2': big['gbs'], 'hbm_frac_level2': big['gbs'] / peak, 'levels': per_level})
    if 'warp_fuse' in legs:
        for C, S, nb in ((64, 96, B), (64, 160, B), (512, 48, 4 * B), (512, 80, 2 * B)):
            for dt in (torch.bfloat16, torch.float32):
                es = 2 if dt == torch.bfloat16 else 4
                feat = Act(torch.rand(nb * FRAMES, S, S, C, generator=g).to(dev).to(dt))
                logits = Act(torch.randn(nb * FRAMES, S, S, C, generator=g).to(dev).to(dt))
                fused = Act.empty(nb, S, S, C, dt, dev)
                for amp in (4.0, 0.8):
                    offs = ((torch.rand(nb * (FRAMES - 1), 2, S, S, generator=g) * 2 - 1) * amp).to(dev)
                    med, mn = timer(lambda: ops.softmax_wsum(feat, logits, fused, FRAMES, offsets=offs))
                    nbytes = nb * ((2 * FRAMES * C * es + (FRAMES - 1) * 2 * 4) * S * S + C * es * S * S)
                    emit({'leg': 'warp_fuse', 'C': C, 'S': S, 'bursts': nb, 'dtype': 'bf16' if es == 2 else 'f32',
                          'flow_px': amp, 'algorithmic_bytes': nbytes, 'us': med * 1e3, 'us_min': mn * 1e3,
                          'hbm_gbs': nbytes / med / 1e6, 'hbm_frac': nbytes / med / 1e6 / peak})
                del feat, logits, fused
    if 'pwc_align' in legs:
        from deep_rawburst_sr_b200.engine import DBSREngine
        from deep_rawburst_sr_b200.models.alignment.pwcnet import PWCNet
        torch.manual_seed(0)
        sd = {'encoder.alignment_net.' + k: v for k, v in PWCNet(load_pretrained=False).state_dict().items()}
        for prec in ('bf16', 'fp32'):
            eng = DBSREngine(sd, dev, precision=prec, parts=('pwc',))
            for S in (96, 160):
                nb = B if prec == 'bf16' else max(1, B // 4)
                burst = torch.rand(nb, FRAMES, 4, S, S, generator=g).to(dev)
                ws = eng.workspace(('micro', nb, S))
                enc_in = eng._buf(ws, 'enc_in', nb * FRAMES, S, S, 8, eng.act_dtype)
                offsets = torch.empty((nb * (FRAMES - 1), 2, S, S), dtype=torch.float32, device=dev)

                def run():
                    eng.prep_and_align(ws, burst, enc_in, offsets)
                n0 = eng.launches
                med, mn = timer(run)
                emit({'leg': 'pwc_align', 'frame': S, 'bursts': nb, 'pairs': nb * (FRAMES - 1), 'precision': prec,
                      'ms': med, 'ms_min': mn, 'pairs_per_s': nb * (FRAMES - 1) / med * 1e3,
                      'launches': (eng.launches - n0) // (args.iters + args.warmup) + 1,
                      'finite': bool(torch.isfinite(offsets).all().item())})
                del burst
                eng._ws.clear()
    if 'sca' in legs:
        # the BurstSR metric path (PWC-Net at the 640^2 output resolution + two warps + colour matching)
        from deep_rawburst_sr_b200.models.alignment.pwcnet import PWCNet
        from deep_rawburst_sr_b200.models.loss.spatial_color_alignment import SpatialColorAlignment
        torch.manual_seed(0)
        for prec in ('bf16', 'fp32'):
            nb = 4 if prec == 'bf16' else 1
            pwc = PWCNet(load_pretrained=False).to(dev).eval().set_precision(prec)
            sca = SpatialColorAlignment(pwc, sr_factor=4)
            sca.to(dev)
            gt = torch.rand(nb, 3, 640, 640, generator=g).to(dev)
            pred = (gt + 0.02 * torch.randn(nb, 3, 640, 640, generator=g).to(dev)).clamp(0, 1)
            burst = torch.rand(nb, 14, 4, 80, 80, generator=g).to(dev)
            med, mn = timer(lambda: sca(pred, gt, burst))
            emit({'leg': 'sca', 'images': nb, 'size': 640, 'alignment_net_precision': prec, 'ms': med, 'ms_min': mn,
                  'images_per_s': nb / med * 1e3})
    if 'metrics' in legs:
        # SSIM / MS-SSIM / PSNR of a batch of predictions (configs[1] / [2] output sizes), one launch
        # chain per call, no host synchronisation.  SSIM: both images read once (8 B per pixel-channel) and 5 moments x
        # 2 x 11 taps of separable window = 110 FMA per window position (+ halo) -> FP32-FMA-bound, reported as GFLOP/s of
        # the algorithmic 2 x 110 FLOP per position next to the HBM figure; PSNR: 8 B per pixel-channel, HBM-bound.
        from deep_rawburst_sr_b200.models.loss import msssim as ms
        from deep_rawburst_sr_b200.models.loss.image_quality_v2 import PSNR
        for nb, S in ((32, 384), (16, 640)):
            gt = torch.rand(nb, 3, S, S, generator=g).to(dev)
            pred = (gt + 0.02 * torch.randn(nb, 3, S, S, generator=g).to(dev)).clamp(0, 1)
            px = nb * 3 * S * S
            pos = nb * 3 * (S - 10) * (S - 10)
            med, mn = timer(lambda: ms.ssim(pred, gt, size_average=False))
            emit({'leg': 'metrics', 'op': 'ssim', 'images': nb, 'size': S, 'ms': med, 'ms_min': mn, 'images_per_s': nb / med * 1e3,
                  'hbm_gbs': 8 * px / med / 1e6, 'hbm_frac': 8 * px / med / 1e6 / peak, 'fp32_gflops': 220 * pos / med / 1e6,
                  'launches': 3})
            med, mn = timer(lambda: ms.ssim(pred, gt, size_average=False, val_range=1.0))
            emit({'leg': 'metrics', 'op': 'ssim_val_range_given', 'images': nb, 'size': S, 'ms': med, 'ms_min': mn,
                  'images_per_s': nb / med * 1e3, 'fp32_gflops': 220 * pos / med / 1e6, 'launches': 2})
            med, mn = timer(lambda: ms.msssim(pred, gt))
            emit({'leg': 'metrics', 'op': 'msssim', 'images': nb, 'size': S, 'ms': med, 'ms_min': mn, 'images_per_s': nb / med * 1e3})
            m = PSNR(boundary_ignore=40)
            med, mn = timer(lambda: m.psnr_per_image(pred, gt))
            cpx = nb * 3 * (S - 80) * (S - 80)
            emit({'leg': 'metrics', 'op': 'psnr_per_image', 'images': nb, 'size': S, 'boundary_ignore': 40, 'ms': med, 'ms_min': mn,
                  'images_per_s': nb / med * 1e3, 'hbm_gbs': 8 * cpx / med / 1e6, 'hbm_frac': 8 * cpx / med / 1e6 / peak})
            del gt, pred
    if 'eval' in legs:
        # the SyntheticBurst scoring protocol (forward -> 14-bit quantisation -> PSNR + SSIM with
        # boundary_ignore 40 -> mean), batched: host bursts / ground truths in, one host read of the report at the end
        import time
        from deep_rawburst_sr_b200.evaluation.synburst.compute_score import TensorBurstSet, score_dataset
        from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic
        torch.manual_seed(0)
        net = dbsrnet_default_synthetic().to(dev).eval().set_precision('bf16')
        net.use_cuda_graph = True
        n = 256
        data = TensorBurstSet(torch.rand(n, FRAMES, 4, 48, 48, generator=g).pin_memory(), torch.rand(n, 3, 384, 384, generator=g).pin_memory())
        score_dataset(net, data, batch_size=32, device=dev)          # warm-up: graph capture, workspaces
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rep = score_dataset(net, data, batch_size=32, device=dev)
        dt = time.perf_counter() - t0                                 # the final report read synchronises
        emit({'leg': 'eval', 'bursts': n, 'batch': 32, 'size': 48, 'precision': 'bf16', 's': dt, 'bursts_per_s': n / dt, 'dataset': 'contiguous pinned tensors',
              'report': rep, 'timing': 'host wall clock around score_dataset (ends with the host read of the report)'})
        items = [data[i] for i in range(n)]          # per-item datasets: stacked into pinned double buffers by the loop
        score_dataset(net, items, batch_size=32, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rep = score_dataset(net, items, batch_size=32, device=dev)
        dt = time.perf_counter() - t0
        emit({'leg': 'eval', 'bursts': n, 'batch': 32, 'size': 48, 'precision': 'bf16', 's': dt, 'bursts_per_s': n / dt, 'dataset': 'per-item list',
              'report': rep, 'timing': 'host wall clock around score_dataset (ends with the host read of the report)'})
    if 'generator' in legs:
        # the synthetic burst generator (default_synthetic.py settings) -- the three kernels alone with the
        # frame transforms prepared once (CUDA events), and `rgb2rawburst` as called (host RNG + sampling per burst, wall clock)
        import time
        from deep_rawburst_sr_b200.data import camera_pipeline as cp
        from deep_rawburst_sr_b200.data import synthetic_burst_generation as G
        params = {'max_translation': 24.0, 'max_rotation': 1.0, 'max_shear': 0.0, 'max_scale': 0.0, 'border_crop': 24}
        img = torch.rand(3, 432, 432, generator=g).to(dev)
        rgb2cam, gains = cp.random_ccm(), cp.random_gains()
        t_mats = G.sample_transforms((432, 432), FRAMES, 4, params)
        z = torch.randn(FRAMES, 4, 48, 48, generator=g).to(dev)

        def kernels_only():
            lin = cp.unprocess(img, rgb2cam, *gains)
            rgb, _ = G.lrburst_from_transforms(lin, t_mats, 4, 24, normalize=True)
            return cp.mosaic_add_noise(rgb, 0.004, 0.0003, z)
        med, mn = timer(kernels_only)
        emit({'leg': 'generator', 'op': 'unprocess + single2lrburst + mosaic_noise (transforms given; includes the per-call upload of '
                                        'the 14 matrices)', 'ms': med, 'ms_min': mn, 'bursts_per_s': 1e3 / med})
        for _ in range(3):
            G.rgb2rawburst(img, FRAMES, 4, dict(params))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            G.rgb2rawburst(img, FRAMES, 4, dict(params))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        emit({'leg': 'generator', 'op': 'rgb2rawburst as called (host RNG, transform sampling, noise draw + upload)', 'ms': dt * 1e3,
              'bursts_per_s': 1.0 / dt, 'timing': 'host wall clock over 50 bursts'})
        # batched front end: parameters of 32 bursts sampled up front, one upload per table, three launches per batch
        imgs = torch.rand(32, 3, 432, 432, generator=g).to(dev)
        for mode in ('device', 'host'):
            for _ in range(2):
                G.rgb2rawburst_batch(imgs, FRAMES, 4, dict(params), noise=mode)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                out = G.rgb2rawburst_batch(imgs, FRAMES, 4, dict(params), noise=mode)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / (reps * 32)
            emit({'leg': 'generator', 'op': f'rgb2rawburst_batch, 32 bursts per call, noise={mode!r}', 'ms_per_burst': dt * 1e3,
                  'bursts_per_s': 1.0 / dt, 'timing': f'host wall clock over {reps} batches of 32', 'burst_shape': list(out[0].shape[1:])})
    if args.out:
        with open(args.out, 'w') as f:
            for d in lines:
                f.write(json.dumps(d) + '\n')


if __name__ == '__main__':
    main()
