#!/bin/bash
# Round-2 GPU session 12: device L-BFGS mode search, C3 overflow fix, batched test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_batched.py -q -x > gpurun_out/s14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s14_pytest.log
timeout 600 python bench.py --workload C3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/s14_bench_c3.log 2>&1
timeout 600 python scripts/prof_init.py C4 > gpurun_out/s14_prof_init.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s14_bench_c4.log 2>&1
tail -25 gpurun_out/s14_pytest.log; grep '^{' gpurun_out/s14_bench_c3.log | tail -1 | cut -c1-200; head -12 gpurun_out/s14_prof_init.log; grep '^{' gpurun_out/s14_bench_c4.log | tail -1 | cut -c1-200
