#!/bin/bash
# DRAM traffic + duration of every launch of one eager forward (after the plain run exits 0)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
CMD="python bench.py --one-forward --warmup 1 --batch ${BENCH_BATCH:-32}"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/plain.log; tail -2 gpurun_out/ncu_traffic.log; wc -l gpurun_out/traffic.csv
