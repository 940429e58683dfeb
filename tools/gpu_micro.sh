#!/bin/bash
# GPU-box script: -m gpu suite, config-4 microbench (bench_micro.py), ncu --set full of the two HBM-side kernels
# (softmax_wsum, corr81) after their plain runs exited 0, default bench, ncu launch list of one forward.  Logs -> gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
if [ "${RUN_TESTS:-1}" = "1" ]; then
  echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x) > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
fi
echo "== micro"; timeout 900 python bench_micro.py --out gpurun_out/micro.json > gpurun_out/micro.log 2>&1; tail -2 gpurun_out/micro.log | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/micro.json'):
    d = json.loads(l)
    if d['leg'] == 'warp_fuse':
        print('warp_fuse', d['C'], d['S'], d['dtype'], d['flow_px'], '%.1f us %.0f GB/s (%.2f of peak)' % (d['us'], d['hbm_gbs'], d['hbm_frac']))
    elif d['leg'] == 'corr81':
        print('corr81', d['frame'], d['dtype'], 'all %.1f us %.0f GB/s' % (d['us_all_levels'], d['hbm_gbs_all_levels']),
              ' '.join('L%d:%.1fus' % (x['level'], x['us']) for x in d['levels']))
    elif d['leg'] == 'pwc_align':
        print('pwc_align', d['frame'], d['precision'], '%.2f ms %.0f pairs/s' % (d['ms'], d['pairs_per_s']))
    elif d['leg'] == 'sca':
        print('sca', d['size'], d['alignment_net_precision'], '%.2f ms %.0f images/s' % (d['ms'], d['images_per_s']))
PY
if [ "${RUN_NCU:-1}" = "1" ]; then
  echo "== ncu"
  python tools/corr_one.py 32 32 104 1 > gpurun_out/corr_one.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:corr81_kernel -s 2 -c 1 -f -o gpurun_out/prof_corr81 python tools/corr_one.py 32 32 104 1 > gpurun_out/ncu_corr.log 2>&1
  cat gpurun_out/corr_one.log
  python tools/wsum_one.py 32 > gpurun_out/wsum_one.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:softmax_wsum -s 2 -c 1 -f -o gpurun_out/prof_wsum python tools/wsum_one.py 32 > gpurun_out/ncu_wsum.log 2>&1
  cat gpurun_out/wsum_one.log
fi
if [ "${RUN_BENCH:-1}" = "1" ]; then
  echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-700
fi
if [ "${RUN_LIST:-0}" = "1" ]; then
  CMD="python bench.py --one-forward --warmup 1 --batch 32"
  $CMD > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
  tail -1 gpurun_out/plain.log; wc -l gpurun_out/launches.csv
fi
