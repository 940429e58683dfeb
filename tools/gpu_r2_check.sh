#!/bin/bash
# round-2 validation call: GPU tests, smoke, bench (48x48 x32), per-layer table
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/gputest.log
echo "pytest rc=$?" >> gpurun_out/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 20 --warmup 5 --layers-out gpurun_out/layers.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/gputest.log; cat gpurun_out/smoke.log; head -c 1500 gpurun_out/bench.json
