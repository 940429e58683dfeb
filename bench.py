#!/usr/bin/env python
"""Benchmark of the DBSR burst forward pass (BASELINE.json metric: bursts/sec, 14-frame, 4x SR).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one forward of a batch of synthetic bursts (BASELINE.json configs[1]: 32 bursts of 14x4x48x48 per GPU,
bf16 tensor cores, random-init weights).  Bursts are independent, so ranks shard bursts with no data-path
collective (weak scaling: the per-GPU batch is fixed); NCCL is used only for the barrier / max-over-ranks timing.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAMES = 14
CONV_GFLOP_PER_BURST = {48: 223.28, 80: 644.10}   # SURVEY.md App. B (2*MAC, all 134 conv calls as the reference runs them)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=32, help='bursts per GPU per step')
    ap.add_argument('--size', type=int, default=48, help='packed RAW height = width')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--cpu-baseline-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--layers-out', default='', help='write the per-conv-layer time table of the instrumented pass here')
    ap.add_argument('--one-forward', action='store_true', help='profiling aid: warm up, run ONE eager forward, exit')
    ap.add_argument('--no-graph', action='store_true', help='launch eagerly instead of replaying a CUDA graph')
    ap.add_argument('--no-overlap', action='store_true', help='run PWC-Net and the encoder on ONE stream (default: two, fork / join)')
    ap.add_argument('--no-extra-configs', action='store_true',
                    help='skip the extra BASELINE.json configs reported next to the headline (cfg_80, cfg_256, strong_cfg2, cfg_256_gather)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops': 1400.0, 'src': 'fallback'}


class ClockSampler(threading.Thread):
    """samples SM clocks / clock-event reasons during the timed regions.  NVML is initialised in the constructor (on the
    main thread, BEFORE any timed region: nvmlInit / an nvidia-smi spawn take the driver lock for tens of ms and stall
    kernel submission -- inside a 50-200 ms region that reads as a 1.2-2.4x slowdown); the thread then only issues
    in-process queries (microseconds).  Only samples taken while `active` is set are kept."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.active = False
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.bits = {'hw_slowdown': pynvml.nvmlClocksEventReasonHwSlowdown,
                         'hw_thermal_slowdown': pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                         'sw_thermal_slowdown': pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                         'sw_power_cap': pynvml.nvmlClocksEventReasonSwPowerCap}
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            nv = self.nvml
            while not self.stop_flag:
                if self.active:
                    try:
                        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                        for n, bit in self.bits.items():
                            if r & bit:
                                self.reasons.add(n)
                    except Exception:
                        pass
                time.sleep(0.005)
            return
        # fallback without NVML bindings: nvidia-smi (coarse; perturbs short regions)
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            if self.active:
                try:
                    out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split('\n')[0]
                    f = [v.strip() for v in out.split(',')]
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                    for n, v in zip(names, f[2:]):
                        if v.lower().startswith('active'):
                            self.reasons.add(n)
                except Exception:
                    pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {'sm_mhz': med, 'sm_min_mhz': s[0] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


def cpu_reference_arm(args, rank):
    """--impl reference: the reference algorithm on the host cores (oracle port of the Python reference: the
    reference is a Python toolkit that cannot travel to the GPU box)."""
    if rank != 0:
        return
    from oracle import dbsr_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict(0)
    burst = O.make_burst(0, 1, FRAMES, args.size, args.size)
    t_w = time.perf_counter()
    for _ in range(max(1, min(args.warmup, 3))):
        O.dbsr_forward_fast(burst, sd)
    per = (time.perf_counter() - t_w) / max(1, min(args.warmup, 3))
    # the same number of timed steps as the product arm, unless that would run for more than ~2.5 minutes on this host
    steps = max(1, min(args.steps, int(150.0 / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.dbsr_forward_fast(burst, sd)
    dt = (time.perf_counter() - t0) / steps
    val = 1.0 / dt
    sample = f'1 burst of {FRAMES}x4x{args.size}x{args.size} per step, fp32, {steps} steps'
    print(json.dumps({
        'impl': 'reference', 'metric': 'bursts/sec (14-frame, 4x SR)', 'value': val, 'unit': 'bursts/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # the product arm's config (same workload string); each step is a bounded sample of it: ONE of its bursts
        'config': {'workload': f'DBSRNet forward, {args.batch} bursts/GPU of {FRAMES}x4x{args.size}x{args.size} packed RAW -> '
                               f'3x{8 * args.size}x{8 * args.size}, random-init weights (BASELINE.json configs[1])',
                   'global_batch': args.batch * args.gpus, 'parallelism': 'host CPU threads (torch intra-op), one burst per step',
                   'sample': sample},
        'cpu_baseline': {'value': val, 'unit': 'bursts/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'bursts/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        cpu_reference_arm(args, rank)
        return
    import torch.distributed as dist
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # random-init weights of the reference architecture: the modules' own default initialisation (+ ICNR), as the
    # reference's factory would produce without a checkpoint.  The product arm never touches oracle/.
    from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic
    torch.manual_seed(0)
    net = dbsrnet_default_synthetic()
    net = net.to(dev).eval().set_precision(args.precision)
    net.use_cuda_graph = not args.no_graph     # the ~137 launches of one forward are captured once per shape and replayed
    eng = net.engine(dev)
    eng.overlap_alignment = not args.no_overlap

    B, S = args.batch, args.size
    gen = torch.Generator().manual_seed(1000 + rank)
    host_in = torch.rand(B, FRAMES, 4, S, S, generator=gen).pin_memory()
    dev_in = host_in.to(dev)
    host_out = torch.empty(B, 3, 8 * S, 8 * S).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        from deep_rawburst_sr_b200.sharding import max_over_ranks as _mx
        return _mx(ms, dev)

    if args.one_forward:
        # profiling aid (ncu launch list / --set full): warm up, then exactly ONE eager forward and exit
        net.use_cuda_graph = False
        for _ in range(max(1, args.warmup)):
            net(dev_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        net(dev_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({'one_forward': True, 'launches_per_forward': eng.launches // (max(1, args.warmup) + 1)}))
        return

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident throughput ("value"): K forwards, inputs already in HBM
    for _ in range(args.warmup):
        net(dev_in)
    for _ in range(2):      # untimed: the CUDA graph of this shape is captured on the 2nd call; make sure replays have run
        net(dev_in)
    barrier()
    sampler.active = True
    eng.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        pred, _aux = net(dev_in)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launches
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- N > 1: the same K forwards, each followed by the NCCL all-gather of `pred` over NVLink on a side stream
    #      (SURVEY.md 8e output gather; OutputGatherer double-buffers so the gather of step i overlaps forward i+1)
    gather = None
    if world > 1:
        try:
            from deep_rawburst_sr_b200.sharding import OutputGatherer
            gat = OutputGatherer(B * world, depth=2)
            for _ in range(2):
                pred, _aux = net(dev_in)
                _full, gev = gat.submit(pred)
            barrier()
            e0.record()
            for _ in range(args.steps):
                pred, _aux = net(dev_in)
                _full, gev = gat.submit(pred)
            torch.cuda.current_stream().wait_event(gev)
            e1.record()
            barrier()
            ms_g = max_over_ranks(e0.elapsed_time(e1))
            gather = {'value': world * B * args.steps / (ms_g / 1e3), 'unit': 'bursts/s', 'ms_per_step': ms_g / args.steps,
                      'collective': 'nccl all_gather_into_tensor of pred on a side stream, overlapped with the next forward',
                      'bytes_gathered_per_rank_per_step': int(_full.numel() * 4),
                      'gathered_shape': list(_full.shape)}
        except Exception as ex:      # the gather is the caller's choice, not part of `value`: report, do not fail the bench
            gather = {'error': repr(ex)[:200]}
        barrier()

    # ---- end to end with HOST buffers, every step: H2D of the burst from pinned memory + forward + D2H of pred.
    #      (a) the public host-buffer front end (HostPipeline: the copies of neighbouring steps overlap the compute on
    #          separate streams; all K copies in and K copies out are inside the timed region, drain included)
    from deep_rawburst_sr_b200.pipeline import HostPipeline
    pipe = HostPipeline(net, depth=2)
    for _ in range(2):
        pipe.submit(host_in, host_out)
    pipe.drain()
    barrier()
    e0.record()
    for _ in range(args.steps):
        pipe.submit(host_in, host_out)
    e1.record(pipe.s_out)
    pipe.drain()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    #      (a') the same with net.output_int16: pred leaves as the reference writers' 14-bit int16 (half the D2H bytes)
    net.output_int16 = True
    host_out16 = torch.empty(B, 3, 8 * S, 8 * S, dtype=torch.int16).pin_memory()
    for _ in range(3):
        pipe.submit(host_in, host_out16)
    pipe.drain()
    barrier()
    e0.record()
    for _ in range(args.steps):
        pipe.submit(host_in, host_out16)
    e1.record(pipe.s_out)
    pipe.drain()
    barrier()
    ms_e2e16 = max_over_ranks(e0.elapsed_time(e1))
    e2e16 = world * B * args.steps / (ms_e2e16 / 1e3)
    net.output_int16 = False
    #      (b) the reference's calling pattern, copies serialised with the forward on one stream: net(x.cuda()) ; pred.cpu()
    for _ in range(2):
        p, _ = net(host_in.to(dev, non_blocking=True))
        host_out.copy_(p, non_blocking=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        p, _ = net(host_in.to(dev, non_blocking=True))
        host_out.copy_(p, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e_sync = max_over_ranks(e0.elapsed_time(e1))
    e2e_sync = world * B * args.steps / (ms_e2e_sync / 1e3)
    sampler.active = False
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=3)

    # ---- the other BASELINE.json configurations the driver never asks for, as extra keys of the same line
    extras = {}
    if not args.no_extra_configs and args.precision == 'bf16' and S == 48 and B == 32:
        net.use_cuda_graph = not args.no_graph

        def time_config(b, s, steps, int16=False, gather_total=0):
            """bursts/s of `steps` forwards of b bursts of 14x4xsxs per GPU (CUDA-graph replay, inputs resident); with
            gather_total > 0 every forward is followed by the NCCL all-gather of the (int16) predictions on a side stream"""
            g2 = torch.Generator().manual_seed(2000 + rank)
            x = torch.rand(b, FRAMES, 4, s, s, generator=g2).to(dev)
            net.output_int16 = int16
            gat = None
            if gather_total:
                from deep_rawburst_sr_b200.sharding import OutputGatherer
                gat = OutputGatherer(gather_total, depth=2)
            gev, full = None, None
            for _ in range(3):
                p_, _a = net(x)
                if gat is not None:
                    full, gev = gat.submit(p_)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(steps):
                p_, _a = net(x)
                if gat is not None:
                    full, gev = gat.submit(p_)
            if gev is not None:
                torch.cuda.current_stream().wait_event(gev)
            a1.record()
            barrier()
            ms = max_over_ranks(a0.elapsed_time(a1))
            net.output_int16 = False
            out = {'bursts_per_gpu': b, 'size': s, 'steps': steps, 'ms_per_step': ms / steps,
                   'value': world * b * steps / (ms / 1e3), 'unit': 'bursts/s'}
            if gat is not None:
                out['gathered_shape'] = list(full.shape)
                out['gathered_dtype'] = str(full.dtype).replace('torch.', '')
                out['bytes_gathered_per_rank_per_step'] = int(full.numel() * full.element_size())
            gf80 = CONV_GFLOP_PER_BURST.get(s)
            if gf80:
                out['model_tflops_per_gpu'] = gf80 * out['value'] / 1e3 / world
                out['frac_of_conv_roofline'] = out['model_tflops_per_gpu'] / peaks()['tflops']
            return out

        try:
            if world == 1:
                # BASELINE configs[2] on ONE GPU (16 bursts of 14x4x80x80 -> 3x640x640): the 1-GPU point of its strong-scaling curve
                extras['cfg_80'] = dict(time_config(16, 80, max(5, args.steps // 2)),
                                        workload='BASELINE.json configs[2] on 1 GPU: 16 bursts of 14x4x80x80 -> 3x640x640')
                # BASELINE configs[4], the share of one GPU: one 14x4x256x256 burst -> 3x2048x2048 (untiled: it fits)
                extras['cfg_256'] = dict(time_config(1, 256, max(5, args.steps // 2)),
                                         workload='BASELINE.json configs[4], one rank\'s share: 1 burst of 14x4x256x256 -> 3x2048x2048, '
                                                  'untiled (the whole crop fits one GPU: tiling would only add halo recompute)')
                # small-batch (latency) regime: what the reference's batch-1 evaluation loops and 8-GPU configs[2] run per GPU
                extras['cfg_48_b2'] = time_config(2, 48, max(10, args.steps))
                extras['cfg_48_b1'] = time_config(1, 48, max(10, args.steps))
            elif 16 % world == 0:
                # STRONG scaling of BASELINE configs[2]: the same 16 bursts of 80x80 split over the ranks, int16 predictions
                # (the reference writers' 14-bit form) all-gathered over NVLink behind the next forward
                extras['strong_cfg2'] = dict(time_config(16 // world, 80, max(5, args.steps // 2), int16=True, gather_total=16),
                                             scaling='strong', global_batch=16,
                                             workload=f'BASELINE.json configs[2]: 16 bursts of 14x4x80x80 split {16 // world} per GPU over '
                                                      f'{world} GPUs, int16 output gather (NCCL all_gather on a side stream)')
                extras['strong_cfg2_no_gather'] = dict(time_config(16 // world, 80, max(5, args.steps // 2)), scaling='strong',
                                                       global_batch=16)
                if world == 8:
                    extras['cfg_256_gather'] = dict(time_config(1, 256, max(5, args.steps // 2), int16=True, gather_total=8),
                                                    workload='BASELINE.json configs[4]: 8 bursts of 14x4x256x256 -> 3x2048x2048, one per GPU '
                                                             '(untiled), int16 NCCL output gather')
        except Exception as ex:      # extras never fail the headline measurement
            extras['error'] = repr(ex)[:300]
        barrier()

    # ---- instrumented pass (same K steps, eager launches bracketed by CUDA events on the launching stream): per
    #      kernel-family device time for the roofline; kept out of the timed regions above because creating ~800 events
    #      per step makes the step CPU-bound
    net.use_cuda_graph = False
    eng.flops = {}
    eng.hbm_bytes = {}
    eng.timers = {}
    eng.layer_events = {}
    for _ in range(args.steps):
        net(dev_in)
    torch.cuda.synchronize()
    fam = eng.timer_summary()
    if rank == 0 and args.layers_out:
        rows = eng.layer_summary()
        with open(args.layers_out, 'w') as f:
            for key, family, shape, n, ms, tf in rows:
                f.write(f'{ms / args.steps:9.4f} ms/step {tf:8.1f} TFLOP/s {family:12s} n,h,w,cin,cout={shape} x{n // args.steps} {key}\n')
    eng.layer_events = None
    flops = dict(eng.flops)
    alg_bytes = dict(eng.hbm_bytes)
    eng.timers = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    # dominant kernel family by measured device time inside the timed region.  The tcgen05 implicit-GEMM convolutions are ONE
    # family for the roofline whether a layer runs as a single conv launch (conv_tc_kernel) or as a fused residual block
    # (resblock32_tc_kernel: two convs per launch) -- moving the slow N = 32 layers into their own row must not flatter it.
    tensor_fams = [k for k in ('conv_tc', 'resblock_tc') if k in fam]
    if tensor_fams:
        fam_tc = (sum(fam[k][0] for k in tensor_fams), sum(fam[k][1] for k in tensor_fams))
        flops['tensor_conv'] = sum(flops.get(k, 0) for k in tensor_fams)
        alg_bytes['tensor_conv'] = sum(alg_bytes.get(k, 0) for k in tensor_fams)
    rank_by = {k: v for k, v in fam.items() if k not in tensor_fams}
    if tensor_fams:
        rank_by['tensor_conv'] = fam_tc
    dom = max(rank_by.items(), key=lambda kv: kv[1][0])[0] if rank_by else None
    roofline = None
    if dom in ('tensor_conv', 'conv_direct'):
        ms, n = rank_by[dom]
        ach = flops.get(dom, 0) / (ms / 1e3) / 1e12
        # measured DRAM traffic of the same launches (ncu dram__bytes_read/write.sum of one forward at this config,
        # committed under profiles/ by tools/gpu_ncu_traffic.sh + tools/traffic_table.py), per launch like `achieved`
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, 'profiles', f'r02_traffic_b{B}.json')
        if S == 48 and args.precision == 'bf16' and os.path.exists(tpath):
            tfile = json.load(open(tpath))
            # the capture must describe THIS launch sequence: same kernel families, same launches per family per forward
            mine = {k: int(round(v[1] / args.steps)) for k, v in fam.items()}
            theirs = {k: int(v['launches']) for k, v in tfile['families'].items()}
            if mine == theirs:
                tfs = [tfile['families'][k] for k in (tensor_fams if dom == 'tensor_conv' else [dom])]
                traffic = sum(t['dram_read_bytes'] + t['dram_write_bytes'] for t in tfs) / sum(t['launches'] for t in tfs)
                traffic_note = f'ncu dram__bytes_read+write.sum of one forward at this config ({os.path.basename(tpath)}), per launch'
            else:
                traffic_note = (f'{os.path.basename(tpath)} does not match this run (launches per family {theirs} vs {mine}): '
                                'regenerate with tools/gpu_ncu_traffic.sh + tools/traffic_table.py')
        roofline = {'kernel': 'conv_tc_kernel + resblock32_tc_kernel (tcgen05 implicit-GEMM convolutions)' if dom == 'tensor_conv' else dom,
                    'bound': 'tensor', 'achieved': ach, 'peak': pk['tflops'], 'unit': 'TFLOP/s',
                    'frac': ach / pk['tflops'], 'traffic': traffic, 'traffic_source': traffic_note, 'peak_source': pk['src'],
                    'launches': n, 'kernel_ms_per_step': ms / args.steps,
                    'flops_per_launch': flops.get(dom, 0) / max(n, 1), 'us_per_launch': ms * 1e3 / max(n, 1),
                    'algorithmic_bytes_per_launch': alg_bytes.get(dom, 0) / max(n, 1),
                    'note': 'family of ~107 implicit-GEMM conv launches per step (Cout 2..512; the four HR residual blocks are one '
                            'fused launch each); SS-mode tcgen05 operand fetch (128 B/clk/SM shared memory) bounds N<128 layers '
                            'below the tensor peak, see DESIGN.md 4.1'}
    elif dom is not None:
        ms, n = fam[dom]
        roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': None, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': None,
                    'traffic': None, 'peak_source': pk['src'], 'launches': n, 'kernel_ms_per_step': ms / args.steps}
    # algorithmic bytes per step of the two HBM-bound fusion kernels (SURVEY.md 8d formulas, element size s)
    es = 2 if args.precision == 'bf16' else 4
    C, HW = 512, S * S
    hbm_bytes = {
        'warp': B * ((FRAMES * C * es) * 2 + (FRAMES - 1) * 2 * 4) * HW,                 # read 14 maps, write 14 maps, flow
        'softmax_wsum': B * (2 * FRAMES * C * es + C * es) * HW,                          # 14 feat + 14 logit maps, fused out
    }
    # cost volume: per pair and pyramid level both feature maps in, 81 channels out
    from deep_rawburst_sr_b200.engine import PWC_EXT_CH
    hp = (S + 63) // 64 * 64
    hbm_bytes['corr81'] = sum(B * (FRAMES - 1) * (2 * PWC_EXT_CH[l] + 81) * (hp >> l) * (hp >> l) * es for l in (2, 3, 4, 5, 6))
    families = {k: {'ms_per_step': v[0] / args.steps, 'launches_per_step': v[1] / args.steps,
                    'tflops': (flops.get(k, 0) / (v[0] / 1e3) / 1e12) if k in flops and v[0] > 0 else None,
                    'hbm_gbs': (hbm_bytes[k] * args.steps / (v[0] / 1e3) / 1e9) if k in hbm_bytes and v[0] > 0 else None}
                for k, v in fam.items()}
    for k in families:
        if families[k]['hbm_gbs'] is not None:
            families[k]['hbm_frac'] = families[k]['hbm_gbs'] / pk['hbm_gbs']
        if families[k]['tflops'] is not None and k in ('conv_tc', 'resblock_tc'):
            families[k]['tensor_frac'] = families[k]['tflops'] / pk['tflops']

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import dbsr_oracle as O   # cpu_baseline leg only: the reference algorithm on the host cores
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = O.make_state_dict(0)
        b1 = O.make_burst(0, 1, FRAMES, S, S)
        O.dbsr_forward_fast(b1, sd)
        t0 = time.perf_counter()
        n = 0
        while True:
            O.dbsr_forward_fast(b1, sd)
            n += 1
            if time.perf_counter() - t0 > args.cpu_baseline_seconds or n >= 20:
                break
        dt = (time.perf_counter() - t0) / n
        cpu = {'value': 1.0 / dt, 'unit': 'bursts/s', 'cores': cores, 'kind': 'port',
               'sample': f'{n} forwards of 1 burst {FRAMES}x4x{S}x{S}, fp32, torch CPU threads={cores}'}

    gf = CONV_GFLOP_PER_BURST.get(S)
    line = {
        'metric': 'bursts/sec (14-frame, 4x SR)', 'value': value, 'unit': 'bursts/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
        'config': {'workload': f'DBSRNet forward, {B} bursts/GPU of {FRAMES}x4x{S}x{S} packed RAW -> 3x{8 * S}x{8 * S}, '
                               f'random-init weights (BASELINE.json configs[1])',
                   'global_batch': B * world, 'parallelism': f'burst-sharded x{world}, no data-path collective',
                   'l2': 'per-step activation working set (~GBs) >> 126 MB L2; no explicit flush'},
        'e2e': {'value': e2e, 'unit': 'bursts/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': host_in.numel() * 4, 'd2h_bytes_per_step': host_out.numel() * 4,
                'api': 'deep_rawburst_sr_b200.pipeline.HostPipeline.submit(host_in, host_out)',
                'int16_output': {'value': e2e16, 'ms_per_step': ms_e2e16 / args.steps, 'd2h_bytes_per_step': host_out16.numel() * 2,
                                 'api': 'net.output_int16 = True ; HostPipeline.submit(host_in, host_out_int16)'},
                'serialised_copies': {'value': e2e_sync, 'ms_per_step': ms_e2e_sync / args.steps,
                                      'api': 'net(host_in.to(device)) ; host_out.copy_(pred)'}},
        'gpu_launches': launches, 'cuda_graph': not args.no_graph, 'alignment_encoder_overlap': not args.no_overlap, 'output_gather': gather,
        'roofline': roofline,
        'kernel_families': families,
        'cpu_baseline': cpu,
        'clocks': sampler.summary(),
        'model_tflops': (gf * value / 1e3) if gf else None,
        'extra_configs': extras or None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
