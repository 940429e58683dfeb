#!/bin/bash
# stability: the whole GPU suite three times, then a long bench run (races / barrier protocol bugs show up as traps or mismatches)
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; done | tee gpurun_out/soak_tests.log
python bench.py --steps 300 --warmup 5 --no-cpu-baseline 2>gpurun_out/soak.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('300 steps: value %.0f ms %.3f e2e %.0f clk %s frac %.3f traffic %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic']))
open('gpurun_out/bench_soak.json','w').write(json.dumps(d))"
tail -3 gpurun_out/soak.err
