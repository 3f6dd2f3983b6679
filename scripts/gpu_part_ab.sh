#!/bin/bash
# A/B of the sliced kernel's work partition (BB_OPT_SELL_PARTITION: 1 equal-cost, 2 slab-aligned, 0 automatic) and of the
# per-slice cost of its cost model (BB_OPT_SELL_SLICE_COST), after the parity tests of the design-matrix products.
mkdir -p gpurun_out
export BENCH_VALUED=0
timeout 900 python -m pytest tests/test_gpu_design.py tests/test_gpu_reference_suite.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -4 gpurun_out/q_pytest.log
run() {  # name workload extra-env...
    local name=$1 wl=$2; shift 2
    env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_$name.log 2>&1
    echo "$name rc=$? $(python - gpurun_out/q_$name.log <<'P'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); o=d['roofline']['other']
    print('it/s %.2f e2e %.2f ms/step %.3f dot %.1f/%.1f tdot %.1f/%.1f us cold/warm' % (d['value'], d['e2e']['value'], d['ms_per_step'],
          1e3*o['spmv_dot_ms'], 1e3*o['spmv_dot_ms_warm'], 1e3*o['spmv_tdot_ms'], 1e3*o['spmv_tdot_ms_warm']))
P
)"
}
run shard8_part1 C4shard8 BB_OPT_SELL_PARTITION=1
run shard8_part0 C4shard8 BB_OPT_SELL_PARTITION=0
run shard8_part0_sc6 C4shard8 BB_OPT_SELL_PARTITION=0 BB_OPT_SELL_SLICE_COST=6
run shard8_part0_sc10 C4shard8 BB_OPT_SELL_PARTITION=0 BB_OPT_SELL_SLICE_COST=10
run c4_part1 C4 BB_OPT_SELL_PARTITION=1
run c4_part0 C4 BB_OPT_SELL_PARTITION=0
run c4_part0_sc6 C4 BB_OPT_SELL_PARTITION=0 BB_OPT_SELL_SLICE_COST=6
run c4_part0_sc10 C4 BB_OPT_SELL_PARTITION=0 BB_OPT_SELL_SLICE_COST=10
run c3_part1 C3 BB_OPT_SELL_PARTITION=1
run c3_part0 C3 BB_OPT_SELL_PARTITION=0
for w in C4shard8 C4; do timeout 250 python scripts/spmv_timeline.py $w > gpurun_out/q_timeline_$w.log 2>&1; done
grep -E "^==|CTA end|sections" gpurun_out/q_timeline_C4shard8.log gpurun_out/q_timeline_C4.log
