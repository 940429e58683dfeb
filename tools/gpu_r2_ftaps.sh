#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/gputest.log
for v in 1 0 1 0; do
 for b in 1 2 32; do
  DBSR_NO_FLOW_TAPS=$v python bench.py --steps 30 --warmup 5 --batch $b --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[no_flow_taps=$v] B=$b value %.0f ms %.3f e2e %.0f launches/step %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']/30))"
 done
done
