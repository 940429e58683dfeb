#!/bin/bash
# round-2: TMA-staged warp-fuse kernel: tests, A/B timing, ncu capture; small-batch bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gputest.log
echo "pytest rc=$?" >> gpurun_out/gputest.log; cat gpurun_out/gputest.log
for b in 32 8; do
  python tools/wsum_one.py $b; DBSR_WSUM_NO_TMA=1 python tools/wsum_one.py $b
done 2>&1 | tee gpurun_out/wsum_ab.log
python bench_micro.py --legs warp_fuse --out gpurun_out/micro_wsum_tma.json > /dev/null 2>&1
DBSR_WSUM_NO_TMA=1 python bench_micro.py --legs warp_fuse --out gpurun_out/micro_wsum_notma.json > /dev/null 2>&1
python tools/wsum_one.py 32 > gpurun_out/wsum_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:softmax_wsum -s 2 -c 1 -f -o gpurun_out/r02_wsum_tma python tools/wsum_one.py 32 > gpurun_out/ncu_wsum.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b32.json 2>gpurun_out/bench_b32.err
python bench.py --steps 50 --warmup 5 --batch 2 --no-cpu-baseline > gpurun_out/bench_b2.json 2>gpurun_out/bench_b2.err
python bench.py --steps 50 --warmup 5 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1.json 2>gpurun_out/bench_b1.err
for f in gpurun_out/bench_b32.json gpurun_out/bench_b2.json gpurun_out/bench_b1.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], 'value %.0f e2e %.0f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()})
PY
done
