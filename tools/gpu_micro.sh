#!/bin/bash
# GPU-box script: full -m gpu suite, config-4 microbench (+ softmax_wsum variants), ncu captures of corr81 / softmax_wsum,
# default bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
if [ "${RUN_TESTS:-1}" = "1" ]; then
  echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x) > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
fi
echo "== micro"; timeout 900 python bench_micro.py --out gpurun_out/micro_v0.json > gpurun_out/micro_v0.log 2>&1; tail -3 gpurun_out/micro_v0.log | cut -c1-400
for v in ${WS_VARIANTS:-1 2 3 4}; do
  DBSR_WS_VARIANT=$v timeout 600 python bench_micro.py --legs warp_fuse --out gpurun_out/micro_v$v.json > gpurun_out/micro_v$v.log 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/micro_v*.json')):
    for l in open(f):
        d = json.loads(l)
        if d['leg'] == 'warp_fuse' and d['dtype'] == 'bf16':
            print(f, d['C'], d['S'], d['flow_px'], '%.1f us %.0f GB/s' % (d['us'], d['hbm_gbs']))
PY
if [ "${RUN_NCU:-1}" = "1" ]; then
  echo "== ncu"
  python tools/corr_one.py 32 32 104 1 > gpurun_out/corr_one.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:corr81_kernel -s 2 -c 1 -f -o gpurun_out/prof_corr81 python tools/corr_one.py 32 32 104 1 > gpurun_out/ncu_corr.log 2>&1
  cat gpurun_out/corr_one.log
  for v in ${NCU_WS_VARIANTS:-0 1}; do
    DBSR_WS_VARIANT=$v python tools/wsum_one.py 32 > gpurun_out/wsum_one_v$v.log 2>&1 && \
    DBSR_WS_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:softmax_wsum -s 2 -c 1 -f -o gpurun_out/prof_wsum_v$v python tools/wsum_one.py 32 > gpurun_out/ncu_wsum_v$v.log 2>&1
    cat gpurun_out/wsum_one_v$v.log
  done
fi
if [ "${RUN_BENCH:-1}" = "1" ]; then
  echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-700
fi
