#!/bin/bash
# Round-2 GPU session 3: n_iter diagnostic, streaming-layout micro-benchmark, fused P-side kernel (tests + bench)
mkdir -p gpurun_out
timeout 300 python scripts/diag_niter.py > gpurun_out/s6_diag.log 2>&1
timeout 120 experimental/_build/stream_bench 256 > gpurun_out/s6_stream.log 2>&1
timeout 900 python -m pytest tests/test_gpu_cg.py tests/test_gpu_gibbs.py tests/test_gpu_multi.py -q -rs > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s6_bench_n1.log 2>&1
BB_OPT_CG_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s6_bench_n1_unfused.log 2>&1
timeout 600 python bench.py --workload C4shard8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s6_bench_shard8.log 2>&1
cat gpurun_out/s6_diag.log gpurun_out/s6_stream.log; tail -15 gpurun_out/s6_pytest.log; for f in s6_bench_n1 s6_bench_n1_unfused s6_bench_shard8; do tail -1 gpurun_out/$f.log | cut -c1-200; done
