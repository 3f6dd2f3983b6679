#!/bin/bash
# Round-2 GPU session (8 GPUs): bench N=8, fused two-shot vs unfused NCCL; C5 on 8 GPUs
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s12_bench_n8.log 2>&1
BB_OPT_CG_FUSED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s12_bench_n8_unfused.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --workload C5 --steps 5 --warmup 3 > gpurun_out/s12_bench_c5_n8.log 2>&1
for f in s12_bench_n8 s12_bench_n8_unfused s12_bench_c5_n8; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-400; tail -3 gpurun_out/$f.log | cut -c1-300; done
