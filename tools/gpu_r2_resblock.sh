#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "resblock32" 2>&1 | tail -15 | tee gpurun_out/resblock_test.log
python tools/resblock_one.py 32 48 two 2>&1 | tail -1 | tee gpurun_out/resblock_ab.log
timeout 120 python tools/resblock_one.py 32 48 fused 2>&1 | tail -1 | tee -a gpurun_out/resblock_ab.log
timeout 120 python tools/resblock_one.py 16 80 fused 2>&1 | tail -1 | tee -a gpurun_out/resblock_ab.log
timeout 120 python tools/resblock_one.py 2 48 fused 2>&1 | tail -1 | tee -a gpurun_out/resblock_ab.log
python tools/wsum_one.py 32; DBSR_WSUM_TMA=0 python tools/wsum_one.py 32; DBSR_WSUM_TMA=2 python tools/wsum_one.py 32
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gputest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers-out gpurun_out/layers_r2.txt > gpurun_out/bench_b32.json 2>gpurun_out/bench_b32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b32.json').read().strip().split('\n')[-1])
print('value %.0f e2e %.0f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()})
PY
