#!/bin/bash
# bench lines at N GPUs of one box (N = first argument); C5 too when the second argument is "c5"; the multi-GPU parity test at N = 2
N=$1
mkdir -p gpurun_out
PORT=$((29500 + N))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m_bench_c4_n$N.log 2>&1
echo "C4 N=$N rc=$? $(grep '^{' gpurun_out/m_bench_c4_n$N.log | tail -1 | cut -c1-160)"
if [ "$2" == "c5" ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT + 20)) \
        bench.py --gpus $N --workload C5 --steps 10 --warmup 3 > gpurun_out/m_bench_c5_n$N.log 2>&1
    echo "C5 N=$N rc=$? $(grep '^{' gpurun_out/m_bench_c5_n$N.log | tail -1 | cut -c1-160)"
fi
if [ "$N" == "2" ]; then
    timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > gpurun_out/m_pytest_multi.log 2>&1
    echo "pytest multi rc=$?"; tail -3 gpurun_out/m_pytest_multi.log
fi
