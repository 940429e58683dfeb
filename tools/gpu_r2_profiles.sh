#!/bin/bash
# round-2 evidence: launch list + DRAM traffic of one forward, ncu --set full of the three new / changed kernels
mkdir -p gpurun_out
CMD="python bench.py --one-forward --warmup 1 --batch 32"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/plain.log; wc -l gpurun_out/traffic.csv
python tools/resblock_one.py 32 48 fused > gpurun_out/rb_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resblock32 -s 2 -c 1 -f -o gpurun_out/r02_resblock32_fused python tools/resblock_one.py 32 48 fused > gpurun_out/ncu_rb.log 2>&1
cat gpurun_out/rb_plain.log
python tools/wsum_one.py 32 > gpurun_out/wsum_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:softmax_wsum -s 2 -c 1 -f -o gpurun_out/r02_wsum_final python tools/wsum_one.py 32 > gpurun_out/ncu_wsum.log 2>&1
cat gpurun_out/wsum_plain.log
python tools/tc_one.py 128 128 3 448 48 48 1 > gpurun_out/tc_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r02_conv_tc_128x128_res_pair python tools/tc_one.py 128 128 3 448 48 48 1 > gpurun_out/ncu_tc.log 2>&1
cat gpurun_out/tc_plain.log
