#!/bin/bash
# ncu --set full capture of the hot kernels in isolation (tcgen05 conv shapes of the encoder / fusion / decoder + the fused
# warp/softmax/weighted-sum), ONE ncu invocation, after the plain run exits 0
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/multi.sh <<'EOS'
python tools/tc_one.py 64 64 3 448 48 48 1 3 && python tools/tc_one.py 32 32 3 32 384 384 0 3 && python tools/tc_one.py 128 128 3 448 48 48 0 3 && python tools/wsum_one.py 32
EOS
bash /tmp/multi.sh > gpurun_out/plain_multi.log 2>&1 && \
ncu --set full --clock-control none --import-source on --target-processes all -k regex:"conv_tc|softmax_wsum" -s 2 -c 1 -f -o gpurun_out/${TC_NAME:-prof_multi} bash /tmp/multi.sh > gpurun_out/ncu_multi.log 2>&1
cat gpurun_out/plain_multi.log; tail -3 gpurun_out/ncu_multi.log
