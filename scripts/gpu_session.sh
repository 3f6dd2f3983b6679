#!/bin/bash
# Round-2 GPU session 5: dense streaming kernel v2, C4 parity test, C2 bench line
mkdir -p gpurun_out
timeout 600 python scripts/dense_bench.py > gpurun_out/s8_dense.log 2>&1
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_design.py -q -k "c4 or dense" > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log
timeout 900 python bench.py --workload C2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s8_bench_c2.log 2>&1
cat gpurun_out/s8_dense.log; tail -8 gpurun_out/s8_pytest.log; tail -3 gpurun_out/s8_bench_c2.log | cut -c1-1800
