#!/bin/bash
mkdir -p gpurun_out
for c in cta_pair_64x64_res_n67 cta_pair_128x128_res_n70_ragged cta_pair_64x512_n99 cta_pair_192x128_n67 cta_pair_1x1_512x64_n67 cta_pair_128x128_f32_n67; do
  timeout 120 python tests/test_gpu_tc.py $c 2>&1 | tail -2
done | tee gpurun_out/pair_cases.log
if grep -q FAIL gpurun_out/pair_cases.log || ! grep -q OK gpurun_out/pair_cases.log; then echo "PAIR CASES FAILED"; fi
for args in "64 64 3 448 48 48 1" "64 64 3 448 48 48 0" "128 128 3 448 48 48 1" "128 128 3 448 48 48 0" "64 512 3 448 48 48 0" "128 512 3 448 48 48 0" "192 128 3 448 48 48 0" "512 64 1 448 48 48 0"; do
  timeout 120 python tools/tc_one.py $args 2>&1 | tail -1
  DBSR_TC_NO_PAIR=1 timeout 120 python tools/tc_one.py $args 2>&1 | tail -1
done | tee gpurun_out/pair_ab.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gputest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --layers-out gpurun_out/layers_pair.txt > gpurun_out/bench_pair.json 2>gpurun_out/bench_pair.err
DBSR_TC_NO_PAIR=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_nopair.json 2>/dev/null
for f in gpurun_out/bench_pair.json gpurun_out/bench_nopair.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], 'value %.0f e2e %.0f ms %.3f clk %s frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['frac']))
PY
done
