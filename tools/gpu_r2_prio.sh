#!/bin/bash
for b in 1 2 32; do
for g in 0 -1 0 -1; do
  DBSR_SIDE_PRIORITY=$g python bench.py --steps 30 --warmup 5 --batch $b --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('[B=$b side priority $g] value %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
done
