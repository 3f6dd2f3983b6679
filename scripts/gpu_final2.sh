#!/bin/bash
# Round-2 closing 1-GPU session: full parity suite + smoke, bench lines of every workload, launch lists, ncu --set full of the SpMV
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/g_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench_c4.log 2>&1
BENCH_NCU_RANGE=1 BENCH_VALUED=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/g_launches_c4.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --clocks none > gpurun_out/g_launch_c4.log 2>&1
BENCH_NCU_RANGE=1 BENCH_VALUED=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/g_launches_shard8.csv python bench.py --workload C4shard8 --steps 2 --warmup 2 --no-cpu-baseline --clocks none > gpurun_out/g_launch_shard8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -s 4 -c 2 -o gpurun_out/g_sell_c4 -f python scripts/prof_spmv.py big > gpurun_out/g_ncu_sell.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -s 4 -c 2 -o gpurun_out/g_sell_shard8 -f python scripts/prof_spmv.py shard8 > gpurun_out/g_ncu_sell8.log 2>&1
for r in g_sell_c4 g_sell_shard8; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.csv 2>/dev/null; done
timeout 600 python bench.py --workload C1 --steps 1000 --warmup 100 > gpurun_out/g_bench_c1.log 2>&1
timeout 600 python bench.py --workload C3 --steps 50 --warmup 10 > gpurun_out/g_bench_c3.log 2>&1
timeout 900 python bench.py --workload C2 --steps 20 --warmup 5 > gpurun_out/g_bench_c2.log 2>&1
timeout 900 python bench.py --workload C2 --sampler cholesky --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_c2_chol.log 2>&1
timeout 900 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_c5.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 2 ) > gpurun_out/g_ref_c4_short.log 2>&1
tail -6 gpurun_out/g_pytest.log; tail -2 gpurun_out/g_smoke.log
for f in g_bench_c4 g_bench_c1 g_bench_c3 g_bench_c2 g_bench_c2_chol g_bench_c5 g_ref_c4_short; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-200; done
tail -4 gpurun_out/g_ref_c4_short.log | cut -c1-200; tail -2 gpurun_out/g_ncu_sell.log gpurun_out/g_ncu_sell8.log
