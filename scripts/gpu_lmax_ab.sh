#!/bin/bash
# A/B of the fragment length cap of the sliced format (BB_OPT_SELL_LMAX; 0 = automatic) on the N = 8 shard size and on C4
mkdir -p gpurun_out
export BENCH_VALUED=0
for l in 0 128 64 32; do
    BB_OPT_SELL_LMAX=$l timeout 300 python bench.py --workload C4shard8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/l_shard8_lmax$l.log 2>&1
    echo "shard8 lmax=$l rc=$? $(grep '^{' gpurun_out/l_shard8_lmax$l.log | tail -1 | cut -c1-100)"
done
for l in 128 64 32; do
    BB_OPT_SELL_LMAX=$l timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/l_c4_lmax$l.log 2>&1
    echo "C4 lmax=$l rc=$? $(grep '^{' gpurun_out/l_c4_lmax$l.log | tail -1 | cut -c1-100)"
done
BB_OPT_SELL_LMAX=64 timeout 300 python bench.py --workload C3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/l_c3_lmax64.log 2>&1
echo "C3 lmax=64 rc=$? $(grep '^{' gpurun_out/l_c3_lmax64.log | tail -1 | cut -c1-100)"
timeout 300 python bench.py --workload C3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/l_c3_lmax0.log 2>&1
echo "C3 lmax=0 rc=$? $(grep '^{' gpurun_out/l_c3_lmax0.log | tail -1 | cut -c1-100)"
