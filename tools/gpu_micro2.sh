#!/bin/bash
# GPU-box script, second iteration: parity of the new softmax_wsum variants + corr81 epilogue, microbench, bench with graphs, launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
for v in 0 5 6; do
  echo "== wsum parity variant $v"
  DBSR_WS_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -k "softmax_wsum" > gpurun_out/wsum_tests_v$v.log 2>&1; tail -4 gpurun_out/wsum_tests_v$v.log
done
echo "== corr parity"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -k "corr81" > gpurun_out/corr_tests.log 2>&1; tail -4 gpurun_out/corr_tests.log
echo "== micro corr"; timeout 600 python bench_micro.py --legs corr81 --out gpurun_out/micro_corr.json > gpurun_out/micro_corr.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/micro_corr.json'):
    d = json.loads(l)
    print(d['frame'], d['dtype'], 'all %.1f us  %.0f GB/s' % (d['us_all_levels'], d['hbm_gbs_all_levels']), ' '.join('L%d:%.1fus' % (x['level'], x['us']) for x in d['levels']))
PY
for v in 5 6; do
  DBSR_WS_VARIANT=$v timeout 600 python bench_micro.py --legs warp_fuse --out gpurun_out/micro_v$v.json > gpurun_out/micro_v$v.log 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/micro_v[56].json')):
    for l in open(f):
        d = json.loads(l)
        if d['leg'] == 'warp_fuse' and d['dtype'] == 'bf16':
            print(f, d['C'], d['S'], d['flow_px'], '%.1f us %.0f GB/s' % (d['us'], d['hbm_gbs']))
PY
for v in 5 6; do
  echo "== forward parity + bench, variant $v"
  DBSR_WS_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu --tb=short -p no:cacheprovider -k "bf16_path_tolerance or full_size" > gpurun_out/fwd_tests_v$v.log 2>&1; tail -3 gpurun_out/fwd_tests_v$v.log; grep "bf16 path" gpurun_out/fwd_tests_v$v.log | head
  DBSR_WS_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_v$v.log 2>&1; tail -1 gpurun_out/bench_v$v.log | cut -c1-330
done
echo "== bench eager (variant 5)"; DBSR_WS_VARIANT=5 timeout 600 python bench.py --no-cpu-baseline --steps 20 --no-graph > gpurun_out/bench_v5_eager.log 2>&1; tail -1 gpurun_out/bench_v5_eager.log | cut -c1-330
echo "== ncu"
DBSR_WS_VARIANT=5 python tools/wsum_one.py 32 > gpurun_out/wsum_one_v5.log 2>&1 && \
DBSR_WS_VARIANT=5 ncu --set full --clock-control none --import-source on -k regex:softmax_wsum -s 2 -c 1 -f -o gpurun_out/prof_wsum_v5 python tools/wsum_one.py 32 > gpurun_out/ncu_wsum_v5.log 2>&1
cat gpurun_out/wsum_one_v5.log
python tools/corr_one.py 32 32 104 1 > gpurun_out/corr_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:corr81_kernel -s 2 -c 1 -f -o gpurun_out/prof_corr81_v2 python tools/corr_one.py 32 32 104 1 > gpurun_out/ncu_corr.log 2>&1
cat gpurun_out/corr_one.log
echo "== launch list"
CMD="python bench.py --one-forward --warmup 1 --batch 32"
DBSR_WS_VARIANT=5 $CMD > gpurun_out/plain.log 2>&1 && \
DBSR_WS_VARIANT=5 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log; wc -l gpurun_out/launches.csv
