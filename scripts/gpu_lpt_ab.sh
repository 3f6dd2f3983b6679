#!/bin/bash
# A/B of the strip assignment inside a section (BB_OPT_SELL_LPT: 1 longest-slice-first + strip-major layout, 0 contiguous cuts)
mkdir -p gpurun_out
export BENCH_VALUED=0
timeout 900 python -m pytest tests/test_gpu_design.py tests/test_gpu_cg.py -m gpu -x -q > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r_pytest.log; tail -4 gpurun_out/r_pytest.log
run() {
    local name=$1 wl=$2 st=$3 wu=$4; shift 4
    env "$@" timeout 300 python bench.py --workload $wl --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r_$name.log 2>&1
    echo "$name rc=$? $(python - gpurun_out/r_$name.log <<'P'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); o=d['roofline']['other']
    print('it/s %.2f e2e %.2f ms/step %.3f ncg %.1f dot %.1f/%.1f tdot %.1f/%.1f us cold/warm' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['mean_n_cg_iter'],
          1e3*o['spmv_dot_ms'], 1e3*o['spmv_dot_ms_warm'], 1e3*o['spmv_tdot_ms'], 1e3*o['spmv_tdot_ms_warm']))
P
)"
}
for l in 0 1; do
    run shard8_lpt$l C4shard8 20 5 BB_OPT_SELL_LPT=$l
    run c4_lpt$l C4 20 5 BB_OPT_SELL_LPT=$l
    run c3_lpt$l C3 50 10 BB_OPT_SELL_LPT=$l
done
run c4_lpt1_again C4 20 5 BB_OPT_SELL_LPT=1
run c4_lpt0_again C4 20 5 BB_OPT_SELL_LPT=0
for w in C4shard8 C4; do BB_OPT_SELL_LPT=1 timeout 250 python scripts/spmv_timeline.py $w > gpurun_out/r_timeline_$w.log 2>&1; done
grep -E "^==|CTA end|warp-end" gpurun_out/r_timeline_C4shard8.log gpurun_out/r_timeline_C4.log
