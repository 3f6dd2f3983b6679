#!/bin/bash
# Round-2 GPU session 7 (2 GPUs): sharded path tests on two devices, bench N=2 fused two-shot vs unfused NCCL
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s10_smi.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -rs -s > gpurun_out/s10_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/s10_multi.log
timeout 600 python -m pytest tests/test_gpu_cg.py tests/test_gpu_gibbs.py -q -k "fisher or cholesky" > gpurun_out/s10_chol.log 2>&1; echo "chol rc=$?" >> gpurun_out/s10_chol.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s10_bench_n2.log 2>&1
BB_OPT_CG_FUSED=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s10_bench_n2_unfused.log 2>&1
grep -E "PASS|FAIL|passed|failed|rc=" gpurun_out/s10_multi.log | tail -8; tail -3 gpurun_out/s10_chol.log; for f in s10_bench_n2 s10_bench_n2_unfused; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-330; done
