#!/bin/bash
# Round-2 final 1-GPU session: full parity suite, bench lines, launch lists, ncu --set full captures
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_c4.log 2>&1
BENCH_NCU_RANGE=1 BENCH_VALUED=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/f_launches_c4.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --clocks none > gpurun_out/f_launch_c4.log 2>&1
BENCH_NCU_RANGE=1 BENCH_VALUED=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/f_launches_shard8.csv python bench.py --workload C4shard8 --steps 2 --warmup 2 --no-cpu-baseline --clocks none > gpurun_out/f_launch_shard8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -s 4 -c 2 -o gpurun_out/f_sell_c4 -f python scripts/prof_spmv.py big > gpurun_out/f_ncu_sell.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -s 4 -c 2 -o gpurun_out/f_sell_shard8 -f python scripts/prof_spmv.py shard8 > gpurun_out/f_ncu_sell8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_stream -s 8 -c 1 -o gpurun_out/f_dense_c2 -f python scripts/prof_dense.py stream > gpurun_out/f_ncu_dense.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_fisher_syrk -c 1 -o gpurun_out/f_fisher_c2 -f python scripts/prof_dense.py fisher > gpurun_out/f_ncu_fisher.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_batch_ -s 4 -c 2 -o gpurun_out/f_batch_c5 -f python scripts/prof_dense.py batch > gpurun_out/f_ncu_batch.log 2>&1
timeout 600 python bench.py --workload C1 --steps 200 --warmup 20 > gpurun_out/f_bench_c1.log 2>&1
timeout 600 python bench.py --workload C3 --steps 50 --warmup 10 > gpurun_out/f_bench_c3.log 2>&1
timeout 900 python bench.py --workload C2 --steps 20 --warmup 5 > gpurun_out/f_bench_c2.log 2>&1
timeout 600 python bench.py --impl reference --workload C3 --steps 5 --warmup 2 > gpurun_out/f_ref_c3.log 2>&1
tail -6 gpurun_out/f_pytest.log; for f in f_bench_c4 f_bench_c1 f_bench_c3 f_bench_c2 f_ref_c3; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-250; done; tail -2 gpurun_out/f_ncu_sell.log gpurun_out/f_ncu_dense.log gpurun_out/f_ncu_fisher.log gpurun_out/f_ncu_batch.log
