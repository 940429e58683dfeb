#!/bin/bash
mkdir -p gpurun_out
run() { # label env...
  echo "== $1"; shift
  for args in "64 64 3 448 48 48 1" "64 64 3 448 48 48 0" "128 128 3 448 48 48 1" "64 512 3 448 48 48 0" "128 512 3 448 48 48 0" "32 32 3 32 384 384 1" "512 64 1 448 48 48 0"; do
    env "$@" timeout 120 python tools/tc_one.py $args 2>&1 | tail -1
  done
}
{
run "pair=0 acc=2" DBSR_TC_PAIR=0 DBSR_TC_ACC_STAGES=2
run "pair=0 acc=4" DBSR_TC_PAIR=0 DBSR_TC_ACC_STAGES=4
run "pair=2 acc=4" DBSR_TC_PAIR=2 DBSR_TC_ACC_STAGES=4
} | tee gpurun_out/pair_ab2.log
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_forward.py -x -q 2>&1 | tail -3
for cfg in "0 2" "0 4" "1 4" "2 4"; do set -- $cfg
  DBSR_TC_PAIR=$1 DBSR_TC_ACC_STAGES=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('pair=$1 acc=$2: value %.0f ms %.3f e2e %.0f clk %s frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac']))"
done | tee gpurun_out/pair_bench2.log
