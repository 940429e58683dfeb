#!/bin/bash
# quick GPU iteration: forward parity + bench with the per-layer table (+ optional ncu launch list)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/quick_tests.log 2>&1; tail -6 gpurun_out/quick_tests.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --batch ${BENCH_BATCH:-32} --no-cpu-baseline --layers-out gpurun_out/layers.txt > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log | cut -c1-1500
head -40 gpurun_out/layers.txt
