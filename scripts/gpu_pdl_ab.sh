#!/bin/bash
# A/B of programmatic dependent launch (BB_OPT_PDL) and the uniform shared-memory carve-out (BB_OPT_UNIFORM_CARVEOUT)
# on one GPU: correctness subset first, then bench lines on the N = 8 shard size and on C4.
mkdir -p gpurun_out
export BENCH_VALUED=0
timeout 900 python -m pytest tests/test_gpu_cg.py tests/test_gpu_multi.py tests/test_gpu_design.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
for cfg in "1 1" "0 0" "1 0" "0 1"; do
    set -- $cfg
    BB_OPT_PDL=$1 BB_OPT_UNIFORM_CARVEOUT=$2 timeout 300 python bench.py --workload C4shard8 --steps 20 --warmup 5 --no-cpu-baseline \
        > gpurun_out/p_shard8_pdl$1_carve$2.log 2>&1
    echo "shard8 pdl=$1 carve=$2 rc=$? $(grep '^{' gpurun_out/p_shard8_pdl$1_carve$2.log | tail -1 | cut -c1-120)"
done
for cfg in "1 1" "0 0"; do
    set -- $cfg
    BB_OPT_PDL=$1 BB_OPT_UNIFORM_CARVEOUT=$2 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline \
        > gpurun_out/p_c4_pdl$1_carve$2.log 2>&1
    echo "C4 pdl=$1 carve=$2 rc=$? $(grep '^{' gpurun_out/p_c4_pdl$1_carve$2.log | tail -1 | cut -c1-120)"
done
for cfg in "1 1" "0 0"; do
    set -- $cfg
    BB_OPT_PDL=$1 BB_OPT_UNIFORM_CARVEOUT=$2 timeout 300 python bench.py --workload C1 --steps 200 --warmup 20 --no-cpu-baseline \
        > gpurun_out/p_c1_pdl$1_carve$2.log 2>&1
    echo "C1 pdl=$1 carve=$2 rc=$? $(grep '^{' gpurun_out/p_c1_pdl$1_carve$2.log | tail -1 | cut -c1-120)"
done
