#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -x -q -k "wsum or forward or golden or bf16" 2>&1 | tail -4
python tools/wsum_one.py 32; python tools/wsum_one.py 8
for gl in 116 124 132 140 148; do
  echo "enc grid limit $gl"; DBSR_ENC_GRID_LIMIT=$gl python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('  value %.0f ms %.3f e2e %.0f clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz']))"
done
