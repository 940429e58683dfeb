#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
if [ "${SKIP_TESTS:-0}" != "1" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputest.log; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/gather_ragged_check.py 2>&1 | grep -v "^W\|warn" | tail -3 | tee gpurun_out/ragged_gather_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().split('\n')[-1])
print('N=$N value %.0f e2e %.0f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'gather', d.get('output_gather',{}).get('value'))
for k,v in (d.get('extra_configs') or {}).items(): print(' ', k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','bursts_per_gpu','bytes_gathered_per_rank_per_step','frac_of_conv_roofline')} if isinstance(v,dict) else v)
PY
