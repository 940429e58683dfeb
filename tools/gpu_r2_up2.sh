#!/bin/bash
mkdir -p gpurun_out
python tools/up_one.py 32 48
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.0f ms %.3f e2e %.0f clk %s frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac']), {k: round(v['ms_per_step'],3) for k,v in d['kernel_families'].items() if k in ('conv_tc','warp_proj','resblock_tc','softmax_wsum')})"
