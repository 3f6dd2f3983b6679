#!/bin/bash
# A/B of the flag-in-data exchange of the fused P-side kernel (BB_OPT_PSIDE_LL) at N GPUs (first argument); parity tests first at N = 2
N=$1
mkdir -p gpurun_out
if [ "$N" == "2" ]; then
    timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -x > gpurun_out/e_pytest_multi.log 2>&1
    echo "pytest multi rc=$?"; tail -3 gpurun_out/e_pytest_multi.log
fi
for ll in ${LL_SEQ:-1 0 1 0}; do
    PORT=$((29600 + RANDOM % 300))
    BB_OPT_PSIDE_LL=$ll timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_bench_n${N}_ll$ll.log 2>&1
    echo "C4 N=$N ll=$ll rc=$? $(grep '^{' gpurun_out/e_bench_n${N}_ll$ll.log | tail -1 | python -c "
import json,sys
l=sys.stdin.read().strip()
if l:
    d=json.loads(l); print('it/s %.2f e2e %.2f ms/step %.3f ncg %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['mean_n_cg_iter']))
")"
    if [ "$ll" == "1" ]; then cp gpurun_out/e_bench_n${N}_ll1.log gpurun_out/e_bench_n${N}_final.log; fi
done
