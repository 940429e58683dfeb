#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --one-forward --warmup 1 --batch 2"
$CMD > gpurun_out/plain_b2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b2.csv $CMD > gpurun_out/ncu_b2.log 2>&1
tail -1 gpurun_out/plain_b2.log; wc -l gpurun_out/launches_b2.csv
