#!/bin/bash
# Round of GPU checks (run under gpurun): per-kernel parity, tcgen05 conv probes (one process per case so a device
# fault in one variant cannot hide the others), end-to-end parity, smoke, short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1; tail -15 gpurun_out/kernels.log
echo "== tc probes"
: > gpurun_out/tc.log
for c in $(python -c "import sys; sys.path.insert(0,\"tests\"); import test_gpu_tc as t; print(\" \".join(t.TC_CASES))"); do
  timeout 120 python tests/test_gpu_tc.py $c >> gpurun_out/tc.log 2>&1 || echo "TC_CASE $c EXIT=$?" >> gpurun_out/tc.log
done
grep -E "TC_CASE|timeout|rror" gpurun_out/tc.log | head -40
echo "== forward"; timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/forward.log 2>&1; tail -25 gpurun_out/forward.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --batch ${BENCH_BATCH:-8} --layers-out gpurun_out/layers.txt > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
