#!/bin/bash
# pside tuning on the N = 8 shard size (one GPU, no exchange) + batched test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_batched.py -q > gpurun_out/s15_batched.log 2>&1; echo "rc=$?" >> gpurun_out/s15_batched.log
for cfg in "0 -1" "64 -1" "32 -1" "16 -1" "0 0" "32 0"; do
  set -- $cfg
  BB_OPT_PSIDE_CTAS=$1 BB_OPT_PSIDE_FOLD_OVF=$2 BENCH_VALUED=0 timeout 300 python bench.py --workload C4shard8 --steps 30 --warmup 5 --no-cpu-baseline --clocks none > gpurun_out/s15_shard8_$1_$2.log 2>&1
  echo "ctas=$1 fold=$2 $(grep '^{' gpurun_out/s15_shard8_$1_$2.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["gpu_launches"])')"
done
tail -3 gpurun_out/s15_batched.log
