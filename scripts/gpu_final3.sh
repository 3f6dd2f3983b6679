#!/bin/bash
# after the last kernel-side change: bench line of C4, launch lists, ncu --set full of the SpMV (C4 and the N = 8 shard size)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench_c4.log 2>&1
BENCH_NCU_RANGE=1 BENCH_VALUED=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/k_launches_c4.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --clocks none > gpurun_out/k_launch_c4.log 2>&1
BENCH_NCU_RANGE=1 BENCH_VALUED=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/k_launches_shard8.csv python bench.py --workload C4shard8 --steps 2 --warmup 2 --no-cpu-baseline --clocks none > gpurun_out/k_launch_shard8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -s 4 -c 2 -o gpurun_out/k_sell_c4 -f python scripts/prof_spmv.py big > gpurun_out/k_ncu_sell.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -s 4 -c 2 -o gpurun_out/k_sell_shard8 -f python scripts/prof_spmv.py shard8 > gpurun_out/k_ncu_sell8.log 2>&1
for r in k_sell_c4 k_sell_shard8; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.csv 2>/dev/null; done
timeout 300 python bench.py --workload C4shard8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/k_bench_shard8.log 2>&1
grep '^{' gpurun_out/k_bench_c4.log | tail -1 | cut -c1-200; tail -n 2 gpurun_out/k_ncu_sell.log; tail -n 2 gpurun_out/k_ncu_sell8.log; ls -la gpurun_out/k_*
