#!/bin/bash
# 1-GPU round-2 check: tests, smoke, bench with the extra configs, then the ncu evidence (tools/gpu_r2_profiles.sh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --layers-out gpurun_out/layers_r2.txt > gpurun_out/bench_b32.json 2>gpurun_out/bench_b32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b32.json').read().strip().split('\n')[-1])
print('value %.0f e2e %.0f ms %.3f roofline %.3f traffic %s cpu %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']))
print({k: round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()})
for k,v in (d.get('extra_configs') or {}).items(): print(' ', k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','frac_of_conv_roofline')} if isinstance(v,dict) else v)
PY
python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
bash tools/gpu_r2_profiles.sh
