#!/bin/bash
# A/B of the taps-per-stage weight ring (DBSR_TC_BSTAGE_KB: 1 = one tap per stage as before, 24 default, 48)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputest.log
for v in "DBSR_TC_BSTAGE_KB=1" "DBSR_TC_BSTAGE_KB=24" "DBSR_TC_BSTAGE_KB=48"; do
 for b in 1 2 32; do
  env $v python bench.py --steps 30 --warmup 5 --batch $b --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[$v] B=$b value %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
 done
done
