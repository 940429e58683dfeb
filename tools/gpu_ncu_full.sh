#!/bin/bash
# ncu --set full capture of ONE tcgen05 conv launch (shape from $TC_SHAPE), after the plain run exits 0
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
CMD="python tools/tc_one.py ${TC_SHAPE:-64 64 3 448 48 48 1}"
$CMD > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/${TC_NAME:-prof_tc} $CMD > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/plain_full.log; tail -3 gpurun_out/ncu_full.log
