#!/bin/bash
# Round-2 GPU session 6: Cholesky comparator (X'WX DMMA kernel, direct draw), C2 with both samplers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cg.py tests/test_gpu_gibbs.py -q -x -k "fisher or cholesky" > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
timeout 900 python bench.py --workload C2 --sampler cholesky --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench_c2_chol.log 2>&1
tail -25 gpurun_out/s9_pytest.log; tail -4 gpurun_out/s9_bench_c2_chol.log | cut -c1-2500
