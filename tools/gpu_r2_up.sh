#!/bin/bash
mkdir -p gpurun_out
python tools/up_one.py 32 48 > gpurun_out/up_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r02_upsample python tools/up_one.py 32 48 > gpurun_out/ncu_up.log 2>&1
cat gpurun_out/up_plain.log
