#!/bin/bash
for g in 116 124 126 128 132 140; do
  DBSR_ENC_GRID_LIMIT=$g python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[limit $g] value %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
