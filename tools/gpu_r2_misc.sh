#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/micro/mma_rate.bin 2 2>&1 | tee gpurun_out/mma_rate_cta2.txt
echo "mma_rate rc=$?"
timeout 900 python -m pytest tests/test_gpu_camera.py tests/test_gpu_eval.py tests/test_gpu_forward.py -x -q 2>&1 | tail -5 | tee gpurun_out/gputest_misc.log
timeout 300 python bench_micro.py --legs generator --out gpurun_out/micro_generator.json 2>&1 | tail -3
python - <<'PY'
import json
for l in open('gpurun_out/micro_generator.json'):
    d=json.loads(l); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ('op','ms','ms_per_burst','bursts_per_s')})
PY
