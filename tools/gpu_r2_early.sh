#!/bin/bash
# A/B of the early weight fetch (weights before griddepcontrol.wait) and of the early launch_dependents trigger
mkdir -p gpurun_out
for v in "DBSR_NO_EARLY_WEIGHTS=1 DBSR_EARLY_TRIGGER=0" "DBSR_EARLY_TRIGGER=0" "DBSR_EARLY_TRIGGER=2" "DBSR_EARLY_TRIGGER=1" "DBSR_NO_EARLY_WEIGHTS=1 DBSR_EARLY_TRIGGER=0" "DBSR_EARLY_TRIGGER=2"; do
 for b in 1 2 32; do
  env $v python bench.py --steps 30 --warmup 5 --batch $b --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[$v] B=$b value %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
 done
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputest.log
