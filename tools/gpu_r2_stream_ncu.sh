#!/bin/bash
# streamed-weight layer of the PWC-Net pyramid (level 3 of two bursts: 26 pairs of 8x8 maps, 565 -> 128, 3x3):
# one tap per weight-ring stage (DBSR_TC_BSTAGE_KB=1, the round-1 / early round-2 form) against the 3-D tap boxes
mkdir -p gpurun_out
for kb in 1 24; do
  for shape in "565 128 3 26 8 8" "661 32 3 26 2 2" "565 128 3 416 8 8"; do
    echo -n "BSTAGE_KB=$kb NSPLIT=on : "; DBSR_TC_BSTAGE_KB=$kb python tools/tc_one.py $shape 0 20
  done
done
echo -n "BSTAGE_KB=24 NSPLIT=off: "; DBSR_TC_NO_NSPLIT=1 python tools/tc_one.py 565 128 3 26 8 8 0 20
echo -n "BSTAGE_KB=1  NSPLIT=off: "; DBSR_TC_NO_NSPLIT=1 DBSR_TC_BSTAGE_KB=1 python tools/tc_one.py 565 128 3 26 8 8 0 20
DBSR_TC_BSTAGE_KB=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r02_conv_tc_stream_1tap python tools/tc_one.py 565 128 3 26 8 8 0 > gpurun_out/ncu_s1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r02_conv_tc_stream_9tap python tools/tc_one.py 565 128 3 26 8 8 0 > gpurun_out/ncu_s9.log 2>&1
ls -la gpurun_out/r02_conv_tc_stream_*
