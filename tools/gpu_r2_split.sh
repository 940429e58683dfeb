#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputest.log
for v in "" "DBSR_NO_SPLIT_WP0=1"; do
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --layers-out gpurun_out/layers_split_$v.txt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[$v] value %.0f ms %.3f e2e %.0f clk %s frac %.3f traffic %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline']['traffic_source']), {k: round(v['ms_per_step'],3) for k,v in d['kernel_families'].items() if k in ('conv_tc','warp_proj','resblock_tc','softmax_wsum')})"
done
head -12 "gpurun_out/layers_split_.txt"
