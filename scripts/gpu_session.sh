#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_design.py -x -q > gpurun_out/s3_design.log 2>&1; echo "design rc=$?" >> gpurun_out/s3_design.log
timeout 600 python scripts/spmv_ab.py C4 1,1np > gpurun_out/s3_ab_c4.log 2>&1
timeout 300 python scripts/spmv_ab.py C4shard8 1 > gpurun_out/s3_ab_shard8.log 2>&1
timeout 300 python scripts/spmv_ab.py C3 1 > gpurun_out/s3_ab_c3.log 2>&1
timeout 300 python scripts/spmv_ab.py C4shard8 1 valued > gpurun_out/s3_ab_shard8_valued.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/s3_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/s3_multi.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s3_bench_n1.log 2>&1
tail -3 gpurun_out/s3_design.log; cat gpurun_out/s3_ab_c4.log gpurun_out/s3_ab_shard8.log gpurun_out/s3_ab_c3.log gpurun_out/s3_ab_shard8_valued.log; tail -30 gpurun_out/s3_multi.log; tail -2 gpurun_out/s3_bench_n1.log
