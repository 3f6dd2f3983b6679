#!/bin/bash
# Round-2 GPU session 10: batched multi-chain path, device-side init pieces, C5 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -q -x > gpurun_out/s13_batched.log 2>&1; echo "rc=$?" >> gpurun_out/s13_batched.log
timeout 900 python -m pytest tests/test_gpu_design.py tests/test_gpu_gibbs.py -q -k "loglik or empty or api or edge or cholesky or (sparse_products and 0-1-2-2)" > gpurun_out/s13_misc.log 2>&1; echo "rc=$?" >> gpurun_out/s13_misc.log
timeout 300 python scripts/spmv_ab.py C3 1p2,2 > gpurun_out/s13_ab_c3.log 2>&1
timeout 300 python scripts/spmv_ab.py C1 1p2,2 > gpurun_out/s13_ab_c1.log 2>&1
timeout 600 python bench.py --workload C5small --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s13_bench_c5small.log 2>&1
timeout 900 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s13_bench_c5.log 2>&1
timeout 600 python scripts/prof_init.py C4 > gpurun_out/s13_prof_init.log 2>&1
tail -30 gpurun_out/s13_batched.log; tail -8 gpurun_out/s13_misc.log; tail -3 gpurun_out/s13_bench_c5small.log | cut -c1-1500; tail -3 gpurun_out/s13_bench_c5.log | cut -c1-2500; head -50 gpurun_out/s13_prof_init.log; cat gpurun_out/s13_ab_c3.log gpurun_out/s13_ab_c1.log
