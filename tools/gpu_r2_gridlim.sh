#!/bin/bash
# encoder grid cap (SMs left to the concurrent PWC-Net stream) at small and full batch
for b in 1 2 4 32; do
for g in 32 64 96 124; do
  DBSR_ENC_GRID_LIMIT=$g python bench.py --steps 30 --warmup 5 --batch $b --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[B=$b limit $g] value %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
done
