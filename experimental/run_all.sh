#!/bin/bash
# First GPU call of the next round: build and run the three prototypes (each prints PASS/FAIL and its timing).
#   gpurun --timeout 600 -- 'bash experimental/run_all.sh > gpurun_out/experimental.log 2>&1; tail -40 gpurun_out/experimental.log'
# (allreduce_twoshot needs >= 2 GPUs: gpurun --gpus 2 / 8; on one GPU it says so and exits)
cd "$(dirname "$0")/.." || exit 1
NVCC="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo"
mkdir -p /tmp/bbx
$NVCC -o /tmp/bbx/spmv_v7 experimental/spmv_v7.cu || exit 1
$NVCC -o /tmp/bbx/cgc experimental/cg_cluster.cu || exit 1
$NVCC -o /tmp/bbx/ar2 experimental/allreduce_twoshot.cu || exit 1
echo "== v7 SpMV: small pattern-only / valued with 12 slabs / many heads per tile / C4-shard / C4 =="
timeout 120 /tmp/bbx/spmv_v7 20000 3000 0.01 1
timeout 120 /tmp/bbx/spmv_v7 20000 3000 0.01 0 1 x 256
timeout 120 /tmp/bbx/spmv_v7 200000 400 0.01 1
timeout 200 /tmp/bbx/spmv_v7 125000 100000 0.001 1
timeout 400 /tmp/bbx/spmv_v7 1000000 100000 0.001 1
echo "== same, next-tile index prefetch after the tile instead of after the gathers =="
$NVCC -DV7_EARLY_PREFETCH=0 -o /tmp/bbx/spmv_v7_late experimental/spmv_v7.cu && timeout 400 /tmp/bbx/spmv_v7_late 1000000 100000 0.001 1
echo "== cluster-fused CG vector kernel =="
timeout 60 /tmp/bbx/cgc 100001 300 16
echo "== two-shot all-reduce =="
timeout 120 /tmp/bbx/ar2 100001 200
