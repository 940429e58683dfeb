#!/bin/bash
# ncu launch list (per-launch device time, serialised / cold cache: compare SHARES) of one eager forward
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
CMD="python bench.py --one-forward --warmup 1 --batch ${BENCH_BATCH:-32}"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
