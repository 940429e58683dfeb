#!/bin/bash
# profiling round: launch list of one forward + ncu --set full of selected kernels (each after its plain run exits 0)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
CMD="python bench.py --one-forward --warmup 1 --batch 32"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log
CMD="python tools/tc_one.py 32 32 3 32 384 384 1"
$CMD > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/prof_tc_32x32 $CMD > gpurun_out/ncu_a.log 2>&1
cat gpurun_out/plain_a.log
CMD="python tools/wsum_one.py 32"
$CMD > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:softmax_wsum -s 2 -c 1 -f -o gpurun_out/prof_wsum $CMD > gpurun_out/ncu_b.log 2>&1
cat gpurun_out/plain_b.log
CMD="python tools/tc_one.py 64 64 3 448 48 48 1"
$CMD > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/prof_tc_64x64 $CMD > gpurun_out/ncu_c.log 2>&1
cat gpurun_out/plain_c.log
